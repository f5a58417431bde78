"""Where does the host-API time of the fused sampler go? (debug helper)"""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import art_planner_b200 as ap
from art_planner_b200 import synth
m = synth.make_fbm_map(400, 400, 0.05, seed=1, amp=0.6) if hasattr(synth, "make_fbm_map") else None
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
L = synth.make_sampler_layers(m, seed=7)
smp = ap.SE3FromSE2Sampler(chk, L, synth.sampler_params_for(m), seed=1)
nd = 1 << 20
h, lib = chk.handle, chk.handle.lib
def t(f, k=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
out = np.empty((nd, 7)); out[:] = 0
nv = C.c_size_t(0)
print("host api, pretouched pageable out: %.2f ms" % t(lambda: lib.artp_sample_valid(h.h, 1, 0, nd, out.ctypes.data, nd, C.byref(nv))))
pin = torch.empty((nd, 7), dtype=torch.float64).pin_memory()
print("host api, pinned out: %.2f ms" % t(lambda: lib.artp_sample_valid(h.h, 1, 0, nd, C.c_void_p(pin.data_ptr()), nd, C.byref(nv))))
print("python wrapper: %.2f ms" % t(lambda: smp.sampleValidBatch(nd, first=0)))
d_out = torch.empty((nd, 7), dtype=torch.float64, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
print("device api: %.2f ms" % t(lambda: smp.sampleValidDevice(nd, 0, d_out, cnt)))
print("np.empty+touch: %.2f ms" % t(lambda: np.empty((nd, 7)).fill(0)))
src = d_out[: nv.value]
print("torch d2h pageable %d MB: %.2f ms" % (src.numel() * 8 >> 20, t(lambda: src.cpu())))
