"""Live parity against the reference's OWN CUDA kernels, compiled source-unmodified by oracle/ref_build/build_ref.sh
into oracle/_ref/ (shipped to the GPU box by gpurun; skipped when absent).  The reference accumulates with fp32
global atomics in arbitrary order, so agreement is to rounding, not bitwise."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

from cdvslam_b200 import synth, fastba, altcorr
from tests.helpers import to_dev, rel_err

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _load(name):
    hits = glob.glob(os.path.join(REF_DIR, name + "*.so"))
    if not hits:
        pytest.skip("oracle/_ref/%s not built" % name)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("eff_impl", [False, True])
def test_ba_matches_reference_cuda_kernels(eff_impl):
    ref = _load("ref_cuda_ba")
    p = synth.config_c2()
    a, b = to_dev(p), to_dev(p)
    ref.forward(a["poses"], a["patches"], a["intrinsics"], a["target"], a["weight"], a["lmbda"], a["ii"], a["jj"],
                a["kk"], p.M, p.t0, p.t1, 2, eff_impl)
    fastba.BA(b["poses"], b["patches"], b["intrinsics"], b["target"], b["weight"], b["lmbda"], b["ii"], b["jj"],
              b["kk"], p.t0, p.t1, M=p.M, iterations=2, eff_impl=eff_impl)
    torch.cuda.synchronize()
    assert rel_err(b["poses"].cpu().numpy(), a["poses"].cpu().numpy()) < 1e-4
    da, db = a["patches"][0, :, 2].cpu().numpy(), b["patches"][0, :, 2].cpu().numpy()
    assert (np.abs(db - da) / np.abs(da)).max() < 1e-4


def test_reproject_matches_reference_cuda_kernel():
    ref = _load("ref_cuda_ba")
    p = synth.config_c2()
    d = to_dev(p)
    a = ref.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    b = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    assert a.shape == b.shape
    assert (a - b).abs().max().item() < 2e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 4e-3)])
def test_corr_and_patchify_match_reference_cuda_kernels(dtype, tol):
    ref = _load("ref_cuda_corr")
    p = synth.small_problem(seed=1, F=8, M=24, t0=3, lifetime=5)
    gmap, pyr = synth.make_fmaps(p, C=24, n_mem=8)
    d = to_dev(p)
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    g = torch.as_tensor(gmap, device="cuda")[None].to(dtype)
    f2 = torch.as_tensor(pyr[0], device="cuda")[None].to(dtype)
    a, = ref.forward(g, f2, coords, d["kk"], d["jj"], 3)
    b = altcorr.corr(g, f2, coords, d["kk"], d["jj"], 3)
    assert a.shape == b.shape
    assert (a.float() - b.float()).abs().max().item() < tol
    xy = torch.rand(1, 64, 2, device="cuda") * torch.tensor([p.wd + 8.0, p.ht + 8.0], device="cuda") - 4.0
    pa, = ref.patchify_forward(f2[0, :1].contiguous(), xy, 1)
    pb = altcorr.patchify(f2[0, :1].contiguous(), xy, 1, mode="none")
    assert torch.equal(pa, pb)                                        # index gather: bit exact
