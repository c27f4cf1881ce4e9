import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import halo2_prover_b200 as h2b
from halo2_prover_b200 import _ffi
import bench, torch
_ffi.init(0)
k = 14
d = h2b.EvaluationDomain(6, k)
cols = [bench.rand_fr_np(1 << k, 300 + i) for i in range(7)]
def T(f, reps=5):
    f(); t = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t) / reps * 1e3
print("single x7 fresh out", T(lambda: [d.coeff_to_extended(c) for c in cols]))
print("many fresh out     ", T(lambda: d.coeff_to_extended_many(cols)))
outs = [np.zeros((d.extended_len(), 4), dtype=np.uint64) for _ in range(7)]
m = 7
pin = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(a) for a in cols])
pout = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(a) for a in outs])
print("many touched out   ", T(lambda: _ffi.check(_ffi.lib().h2b_coeff_to_extended_many(C.byref(d._d), pin, pout, C.c_size_t(m)))))
print("single touched out ", T(lambda: [_ffi.check(_ffi.lib().h2b_coeff_to_extended(C.byref(d._d), _ffi.u64p(cols[q]), _ffi.u64p(outs[q]))) for q in range(7)]))
din = torch.from_numpy(np.concatenate(cols).view(np.int64)).cuda()
dout = torch.empty((7 << d.extended_k, 4), dtype=torch.int64, device="cuda")
s = torch.cuda.Stream()
def dev():
    _ffi.check(_ffi.lib().h2b_dev_coeff_to_extended_many(C.byref(d._d), C.c_void_p(din.data_ptr()), C.c_void_p(dout.data_ptr()), C.c_size_t(7), C.c_void_p(s.cuda_stream)))
    s.synchronize()
print("dev many           ", T(dev))
def dev1():
    for q in range(7):
        _ffi.check(_ffi.lib().h2b_dev_coeff_to_extended(C.byref(d._d), C.c_void_p(din.data_ptr() + q * (32 << k)), C.c_void_p(dout.data_ptr() + q * (32 << d.extended_k)), C.c_void_p(s.cuda_stream)))
    s.synchronize()
print("dev single x7      ", T(dev1))
ho = torch.empty((7 << d.extended_k, 4), dtype=torch.int64).pin_memory()
def d2h():
    ho.copy_(dout, non_blocking=True); torch.cuda.synchronize()
print("D2H 28 MB pinned   ", T(d2h))
