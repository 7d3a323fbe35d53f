import sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from laplacian import load_L
from filtered_lanczos import lanczos
name=sys.argv[1]
L,dg=load_L(name); n=L.shape[0]
b=2*dg.max()*(1+1e-9); a0=max(dg.min()*n/(n-1)*1.01,b/1024)
rng=np.random.default_rng(1); v0=rng.random(n)-0.5
print(name,'n',n,'b %.1f a0 %.3f'%(b,a0))
for a in (a0, a0/3, a0/10):
    for d in (16,32,64):
        t=time.time()
        r=lanczos(L,a,b,d,v0,400 if d==16 else 150,1e-9,1.0,check_every=2)
        print('  a=%.4f d=%3d: steps %3d matvecs %5d lam2 %s res %s (%.1fs)'%(a,d,r['steps'],r['nmv'],r['lam2'],r['res'],time.time()-t),flush=True)
