"""Time the reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100a) next to ours on the same
B200: fastba.BA on the c2 window and altcorr.corr on the c3 shape.  Prints one JSON line.  Measurement aid only."""
import glob, importlib.util, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np
import torch
from cdvslam_b200 import synth, fastba, altcorr
from tests.helpers import to_dev


def load(name):
    hits = glob.glob(os.path.join(REPO, "oracle", "_ref", name + "*.so"))
    spec = importlib.util.spec_from_file_location(name, hits[0])
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    out = {}
    ref_ba, ref_corr = load("ref_cuda_ba"), load("ref_cuda_corr")
    p = synth.config_c2()
    d = to_dev(p, pad_pose_rows=4096 - 22, pad_patch_rows=(4096 - 22) * 96)
    p0, q0 = d["poses"].clone(), d["patches"].clone()
    def reset(): d["poses"].copy_(p0); d["patches"].copy_(q0)
    def ours(): reset(); fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=96, iterations=2)
    def ref(eff):
        def f(): reset(); ref_ba.forward(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], 96, p.t0, p.t1, 2, eff)
        return f
    out["ba_c2_ms"] = {"ours_api": timeit(ours), "reference_dense": timeit(ref(False)), "reference_eff": timeit(ref(True))}
    if os.environ.get("REF_C4", "1") == "1":
        p4 = synth.config_c4()
        d4 = to_dev(p4, pad_pose_rows=4096 - 1000, pad_patch_rows=(4096 - 1000) * 96)
        a0, b0 = d4["poses"].clone(), d4["patches"].clone()
        def reset4(): d4["poses"].copy_(a0); d4["patches"].copy_(b0)
        def ours4(): reset4(); fastba.BA(d4["poses"], d4["patches"], d4["intrinsics"], d4["target"], d4["weight"], d4["lmbda"], d4["ii"], d4["jj"], d4["kk"], p4.t0, p4.t1, M=96, iterations=2, eff_impl=True)
        def ref4(): reset4(); ref_ba.forward(d4["poses"], d4["patches"], d4["intrinsics"], d4["target"], d4["weight"], d4["lmbda"], d4["ii"], d4["jj"], d4["kk"], 96, p4.t0, p4.t1, 2, True)
        out["ba_c4_ms"] = {"ours_api": timeit(ours4, 5, 2), "reference_eff": timeit(ref4, 3, 1)}
        del d4, a0, b0
        torch.cuda.empty_cache()
    for C, dt in ((24, torch.float16), (128, torch.float32)):
        gmap, pyr = synth.make_fmaps(p, C=C)
        g = torch.as_tensor(gmap, device="cuda")[None].to(dt)
        f0 = torch.as_tensor(pyr[0], device="cuda")[None].to(dt)
        f1 = torch.as_tensor(pyr[1], device="cuda")[None].to(dt)
        coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
        kk, jj = d["kk"], d["jj"]
        def ours_c(): return torch.stack([altcorr.corr(g, f0, coords, kk, jj, 3), altcorr.corr(g, f1, coords / 4, kk, jj, 3)], -1).view(1, len(kk), -1)
        def ours_f(): return altcorr.corr_pyramid2(g, [f0, f1], coords, kk, jj, 3)
        def ref_c(): return torch.stack([ref_corr.forward(g, f0, coords, kk, jj, 3)[0], ref_corr.forward(g, f1, coords / 4, kk, jj, 3)[0]], -1).view(1, len(kk), -1)
        out["corr_c3_C%d_%s_ms" % (C, str(dt).split(".")[-1])] = {"ours_two_calls": timeit(ours_c, 20), "ours_fused": timeit(ours_f, 20), "reference": timeit(ref_c, 20)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
