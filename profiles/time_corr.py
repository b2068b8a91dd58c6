"""Times altcorr.corr_pyramid2 on the c3 shape (c2 graph, C=24 fp16, 120x160 + 30x40, radius 3) for the available
kernels (PCORR_TMA=1 default / PCORR_TMA=0 staged), L2 flushed between calls.  Usage: python profiles/time_corr.py"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
from cdvslam_b200 import synth, fastba, altcorr   # noqa: E402
import bench                                      # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    p = synth.config_c2()
    d = synth.to_torch(p, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
    res = {}
    for C in (24, 32, 128):
        gmap, pyr = synth.make_fmaps(p, C=C)
        g = torch.as_tensor(gmap, device=dev)[None].half()
        f0 = torch.as_tensor(pyr[0], device=dev)[None].half()
        f1 = torch.as_tensor(pyr[1], device=dev)[None].half()
        for mode in ("1", "0"):
            os.environ["PCORR_TMA"] = mode
            fn = lambda: altcorr.corr_pyramid2(g, [f0, f1], coords, d["kk"], d["jj"], 3)
            for _ in range(3):
                fn()
            res["C%d_tma%s_eager_ms" % (C, mode)] = bench.timed_events(fn, 20, before=flush.zero_)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            res["C%d_tma%s_graph_ms" % (C, mode)] = bench.timed_events(gr.replay, 20, before=flush.zero_)
    # fp32, C = 128: TMA tile kernel (3xTF32) vs the staged FFMA kernel
    gmap, pyr = synth.make_fmaps(p, C=128)
    g = torch.as_tensor(gmap, device=dev)[None].float()
    f0 = torch.as_tensor(pyr[0], device=dev)[None].float()
    f1 = torch.as_tensor(pyr[1], device=dev)[None].float()
    for mode in ("1", "0"):
        os.environ["PCORR_TMA"] = mode
        fn = lambda: altcorr.corr_pyramid2(g, [f0, f1], coords, d["kk"], d["jj"], 3)
        for _ in range(3):
            fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        res["C128_f32_tma%s_graph_ms" % mode] = bench.timed_events(gr.replay, 20, before=flush.zero_)
    os.environ["PCORR_TMA"] = "1"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
