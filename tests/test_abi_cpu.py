"""CPU-side checks: the C-ABI library loads and exports every symbol the headers declare; argument validation works
without a GPU (no compute is launched); host logic (neighbors, synthetic generator) matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from cdvslam_b200 import native, synth
from oracle import neighbors_oracle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("pgba.h", "pcorr.h"):
        txt = open(os.path.join(REPO, "include", h)).read()
        names |= set(re.findall(r"\b(p(?:gba|corr)_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    L = native.lib()
    decl = _declared_symbols()
    assert {"pgba_ba_solve", "pgba_ba_solve_batched", "pgba_reproject", "pcorr_forward", "pcorr_patchify_forward",
            "pcorr_backward", "pcorr_patchify_backward", "pgba_ba_linearize_debug"} <= decl
    for name in decl:
        assert hasattr(L, name), name
    assert set(native.SIGNATURES) == decl


def test_workspace_query_and_argument_validation():
    L = native.lib()
    n = ctypes.c_size_t(0)
    assert L.pgba_ba_workspace_bytes(37824, 4096, 4096 * 96, 12, 22, 1, ctypes.byref(n)) == 0
    one = n.value
    assert one > 37824 * 4 and one % 256 == 0
    assert L.pgba_ba_workspace_bytes(37824, 4096, 4096 * 96, 12, 22, 64, ctypes.byref(n)) == 0
    assert one < n.value <= 64 * one and n.value % 256 == 0     # chunk size adapts to the batch
    assert L.pgba_ba_workspace_bytes(-1, 10, 10, 0, 1, 1, ctypes.byref(n)) == -2
    assert L.pgba_ba_workspace_bytes(10, 10, 10, 0, 1, 1, None) == -1
    # NULL pointers / bad shapes are rejected before any CUDA call
    assert L.pgba_ba_solve(None, None, None, None, None, None, None, None, None, 10, 10, 10, 3, 96, 1, 5, 2, 0,
                           None, 0, None) == -1
    assert L.pgba_error_string(-3).decode().startswith("pgba")
    assert L.pcorr_forward(None, None, None, None, None, 1, 5, 5, 5, 8, 10, 10, 3, 3, 0, None, None) == -1
    assert L.pcorr_forward(None, None, None, None, None, 1, 0, 5, 5, 8, 10, 10, 3, 3, 0, None, None) == 0


def test_drop_in_module_surfaces():
    """Same names as the reference pybind tables (ba.cpp:183-188, correlation.cpp:57-63) and python shims."""
    import cuda_ba, cuda_corr
    from cdvslam_b200 import fastba, altcorr
    for f in ("forward", "neighbors", "reproject", "solve_system"):
        assert callable(getattr(cuda_ba, f))
    for f in ("forward", "backward", "patchify_forward", "patchify_backward"):
        assert callable(getattr(cuda_corr, f))
    import inspect
    assert list(inspect.signature(fastba.BA).parameters) == ["poses", "patches", "intrinsics", "target", "weight",
                                                             "lmbda", "ii", "jj", "kk", "t0", "t1", "M", "iterations",
                                                             "eff_impl"]
    assert list(inspect.signature(altcorr.corr).parameters) == ["fmap1", "fmap2", "coords", "ii", "jj", "radius",
                                                                "dropout"]
    assert list(inspect.signature(altcorr.patchify).parameters) == ["net", "coords", "radius", "mode"]


def test_no_cpu_fallback():
    from cdvslam_b200 import fastba
    p = synth.small_problem()
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a), dtype=dt)
    with pytest.raises(RuntimeError, match="CUDA"):
        fastba.BA(t(p.poses)[None], t(p.patches)[None], t(p.intrinsics)[None], t(p.target)[None], t(p.weight)[None],
                  t([1e-4]), t(p.ii, torch.int64), t(p.jj, torch.int64), t(p.kk, torch.int64), p.t0, p.t1, M=p.M,
                  iterations=2)


def test_neighbors_abi_validation_and_oracle_self_check():
    """pgba_neighbors rejects bad arguments before any CUDA call; the oracle's lexsort form equals the literal
    group-then-stable-sort loop of the reference (ba.cpp:66-90)."""
    L = native.lib()
    n = ctypes.c_size_t(0)
    assert L.pgba_neighbors_workspace_bytes(37824, ctypes.byref(n)) == 0 and n.value % 256 == 0 and n.value > 8 * 37824
    assert L.pgba_neighbors_workspace_bytes(-1, ctypes.byref(n)) == -2
    assert L.pgba_neighbors(None, None, 10, None, None, None, 0, None) == -1
    assert L.pgba_neighbors(None, None, 0, None, None, None, 0, None) == 0
    import cuda_ba
    with pytest.raises(RuntimeError, match="CUDA"):
        cuda_ba.neighbors(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64))
    rng = np.random.default_rng(0)
    ii = rng.integers(0, 40, 500); jj = rng.integers(0, 12, 500)
    oi, oj = neighbors_oracle.neighbors(ii, jj)
    ix = np.full(500, -1); jx = np.full(500, -1)
    for g in np.unique(ii):                                   # ba.cpp:79-90, literally
        idx = sorted(np.nonzero(ii == g)[0].tolist(), key=lambda e: jj[e])      # sorted() is stable
        for t, e in enumerate(idx):
            ix[e] = idx[t - 1] if t > 0 else -1
            jx[e] = idx[t + 1] if t + 1 < len(idx) else -1
    np.testing.assert_array_equal(ix, oi)
    np.testing.assert_array_equal(jx, oj)


def test_synthetic_configs_have_the_surveyed_sizes():
    c1, c2 = synth.config_c1(), synth.config_c2()
    assert (c1.E, c1.N, c1.patches.shape[0]) == (9600, 9, 960)
    assert (c2.E, c2.N, c2.patches.shape[0]) == (37824, 10, 2112)
    assert len(set(zip(c2.ii.tolist(), c2.jj.tolist()))) == 394
    ii, jj, kk = synth.global_edges(1000, 96, 200, np.random.default_rng(1241))
    assert len(ii) == 402624
