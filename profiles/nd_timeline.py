"""Per-launch timeline of the reordered global-BA factorisation (config c4, one iteration): needs a -DPGBA_ND_TIMING build
passed through PGBA_LIB (csrc/build_variant.sh ndts -DPGBA_ND_TIMING).  Prints, per panel step, when each launch was
entered / released by its dependency / finished (us, relative to the first)."""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "cdv-slam_b200")]
import numpy as np, torch
from cdvslam_b200 import synth, fastba, native
from tests.helpers import to_dev
p = synth.config_c4()
d = to_dev(p)
p0, q0 = d["poses"].clone(), d["patches"].clone()
def call():
    d["poses"].copy_(p0); d["patches"].copy_(q0)
    fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=1, eff_impl=True)
    torch.cuda.synchronize()
call(); call()
lib = ctypes.CDLL(native.LIB_PATH)
raw = np.zeros(8 * 1024 * 3, np.uint64)
buf = raw.reshape(8 * 1024, 3)
lib.pgba_nd_timestamps(None, 1)
call()
lib.pgba_nd_timestamps(raw.ctypes.data_as(ctypes.c_void_p), 0)
ok = buf[:, 2] > 0
t0 = buf[ok][:, 0].min()
names = ["potf2", "trsm", "syrk", "gather", "backsolve", "border(coop)", "finish", "?"]
rows = []
for i in np.nonzero(ok)[0]:
    kind, mode, idx = i // 1024, (i % 1024) // 512, i % 512
    e, w, x = [(int(v) - int(t0)) / 1e3 for v in buf[i]]
    rows.append((e, names[kind], mode, idx, w, x))
rows.sort()
prev_end = 0.0
for e, n, mode, idx, w, x in rows:
    if mode == 1 and idx > 36: continue
    print("%-6s mode %d step %3d  entered %8.2f  released %8.2f  end %8.2f  body %6.2f  gap after prev end %6.2f" % (n, mode, idx, e, w, x, x - w, w - prev_end))
    prev_end = x

