import sys, ctypes as C; sys.path.insert(0,'.')
import torch
from optika_b200 import _lib
lib=_lib.lib()
for n in (10_000_000, 100_000_000):
    g=C.c_double()
    _lib.check(lib.optk_measure_soa_copy(n, C.byref(g), None))
    print('soa copy', n, g.value, 'GB/s')
a=torch.empty(1<<30, dtype=torch.bfloat16, device='cuda'); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
print('torch copy GB/s', 2*a.numel()*2/ (e0.elapsed_time(e1)*1e-3)/1e9)
