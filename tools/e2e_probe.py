import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from cuzk_b200 import api, lib as cl
api.initialize(0); L = cl.get_lib()
n = 1_000_000
dl = torch.empty((n,4),dtype=torch.int64,device='cuda'); dr=torch.empty_like(dl)
L.cuzk_synth_elements(dl.data_ptr(), n, 1, 0, 1, None); L.cuzk_synth_elements(dr.data_ptr(), n, 2, 0, 1, None)
hl = dl.cpu().pin_memory(); hr = dr.cpu().pin_memory(); ho = torch.empty((n,4),dtype=torch.int64).pin_memory()
for _ in range(3): L.cuzk_poseidon_hash_pairs(hl.data_ptr(), hr.data_ptr(), ho.data_ptr(), n, 1, None)
t0=time.perf_counter()
for _ in range(20): L.cuzk_poseidon_hash_pairs(hl.data_ptr(), hr.data_ptr(), ho.data_ptr(), n, 1, None)
dt=(time.perf_counter()-t0)/20
print(os.environ.get('CUZK_CHUNK_PERCENT','100'), f"{dt*1e3:.3f} ms  {n/dt/1e6:.1f} M/s")
