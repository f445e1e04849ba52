import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from test_gpu_umma import _run
for K,N in [(64,64),(128,64),(64,128)]:
    for p in (1,3):
        got,ref=_run(K,N,p,seed=3)
        e=np.abs(got-ref); 
        import numpy as np
        rng=np.random.default_rng(3); A=rng.standard_normal((128,K)).astype(np.float32); B=rng.standard_normal((N,K)).astype(np.float32)
        f32=np.abs((A@B.T).astype(np.float64)-ref)
        print(K,N,p,"max",e.max(),"rms",np.sqrt((e**2).mean()),"| fp32 numpy max",f32.max(),"rms",np.sqrt((f32**2).mean()))
