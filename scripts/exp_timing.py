"""Read the wait-time counters of an EXP_TIMING build of the backward kernel (build.py -D EXP_TIMING=1).  Diagnostic only."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SELSCAN_B200_LIB"] = os.path.join(ROOT, "mamba-unet_b200/lib/exp_timing.so")
sys.argv = [sys.argv[0], sys.argv[1] if len(sys.argv) > 1 else "S1", "24", "1"]
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
from selscan_b200 import _lib
lib = _lib.load()
buf = (ctypes.c_ulonglong * 8)()
exec(open(os.path.join(ROOT, "scripts/run_scan_once.py")).read())     # warm-up launch
lib.selscan_b200_debug_tw(buf, 1)
ops.launch_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None, t["ck"], True,
               t["du"], t["dd"], dA, dB, dC, dD, None, db)
torch.cuda.synchronize()
lib.selscan_b200_debug_tw(buf, 0)
v = list(buf)
tot = v[3]
print("helper (warp 4 lane 0), fraction of its lifetime: wait tma_full(prep) %.3f  half_full[0] %.3f  half_full[1] %.3f" % (v[0] / tot, v[1] / tot, v[2] / tot))
print("compute (warp 0 lane 0): wait prep_done %.3f  half_free[0] %.3f  half_free[1] %.3f  tma_full %.3f" % (v[4] / tot, v[5] / tot, v[6] / tot, v[7] / tot))
print("cycles per CTA", tot / (b * 4 * (kd // 4) // 64))
