import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from trajopt_grpo_b200 import _lib as L
lib = L.load()
for n in (48, 96, -48, -96, -10048, -10096):
    out = (C.c_longlong * 4)()
    L.check(lib.tg_tmem_probe(L.ctx(torch.device('cuda', 0)), n, out), 'probe')
    print(f"n_mma={n:4d}  ld {out[0]:6d} clk   st {out[1]:6d} clk   mma issue->retire {out[2]:6d} clk  (issue {out[3]})")
