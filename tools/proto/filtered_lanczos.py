import sys, numpy as np, scipy.sparse as sp, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from laplacian import load_L
def cheb_apply(L,x,a,b,d):
    # T_d((b+a-2L)/(b-a)) x
    c=(b+a)/2; e=(b-a)/2
    y0=x; y1=(c*x-L@x)/e
    for k in range(2,d+1):
        y0,y1=y1,2*(c*y1-L@y1)/e-y0
    return y1 if d>=1 else x
def lanczos(L,a,b,d,v0,maxsteps,tol_rel,lam_scale,check_every=2,nev=2):
    n=L.shape[0]
    V=np.zeros((n,maxsteps+1)); al=[];be=[]
    v=v0/np.linalg.norm(v0); V[:,0]=v; nmv=0
    for j in range(maxsteps):
        w=cheb_apply(L,V[:,j],a,b,d); nmv+=d
        h=V[:,:j+1].T@w; w-=V[:,:j+1]@h
        h2=V[:,:j+1].T@w; w-=V[:,:j+1]@h2
        al.append(h[j]+h2[j]); bt=np.linalg.norm(w); be.append(bt); V[:,j+1]=w/bt
        if (j+1)>=6 and (j+1)%check_every==0:
            T=np.diag(al)+np.diag(be[:-1],1)+np.diag(be[:-1],-1)
            th,Y=np.linalg.eigh(T)
            # top nev of B
            X=V[:,:j+1]@Y[:,-nev:]
            # Rayleigh-Ritz on L
            LX=L@X; nmv+=nev
            H=X.T@LX; H=(H+H.T)/2
            mu,Z=np.linalg.eigh(H)
            lam2=mu[-1]; x=X@Z[:,-1]; r=np.linalg.norm(L@x-lam2*x)
            nmv+=0
            if r<=max(tol_rel*abs(lam2),1e-13*b):
                return dict(steps=j+1,nmv=nmv,lam2=lam2,res=r,x=x,theta=th,al=al,be=be,V=V[:,:j+1],Y=Y)
    T=np.diag(al)+np.diag(be[:-1],1)+np.diag(be[:-1],-1)
    th,Y=np.linalg.eigh(T)
    return dict(steps=maxsteps,nmv=nmv,lam2=None,res=None,theta=th,al=al,be=be,V=V[:,:maxsteps],Y=Y)
def ritz_to_lambda(th,a,b,d):
    # invert T_d(t)=th for th>=1 : t=cosh(acosh(th)/d); lambda=(b+a-t(b-a))/2
    th=np.maximum(th,1.0)
    t=np.cosh(np.arccosh(th)/d)
    return (b+a-t*(b-a))/2
if __name__=='__main__':
    name=sys.argv[1]
    L,dg=load_L(name); n=L.shape[0]
    b=2*dg.max()*(1+1e-9); a0=max(dg.min()*n/(n-1)*1.01,b/1024)
    rng=np.random.default_rng(1); v0=rng.random(n)-0.5
    t=time.time()
    base=lanczos(L,a0,b,16,v0,300,1e-9,1.0,check_every=4)
    print(name,'baseline a=%.3f b=%.1f d=16: steps %d matvecs %d lam2 %.10g res %.2e  (%.1fs)'%(a0,b,base['steps'],base['nmv'],base['lam2'],base['res'],time.time()-t))
    for m1 in (12,20):
      for K in (6,10):
        for target in (2.0,3.0):
            t=time.time()
            p1=lanczos(L,a0,b,16,v0,m1,1e-30,1.0,check_every=10**6)
            lam_est=np.sort(ritz_to_lambda(p1['theta'],a0,b,16))
            a1=max(lam_est[min(K,len(lam_est)-1)],b/65536)*1.0
            # degree so that d*sqrt(2a/(b-a)) ~ target
            d1=int(np.clip(round(target/np.sqrt(2*a1/(b-a1))),8,512))
            # start vector: sum of the two top Ritz vectors of phase 1
            x0=p1['V']@p1['Y'][:,-2:].sum(axis=1)
            p2=lanczos(L,a1,b,d1,x0,200,1e-9,1.0,check_every=1)
            tot=p1['nmv']+(p2['nmv'])
            print('  m1=%d K=%d target=%.1f -> a1=%.4f d1=%d : phase2 steps %s matvecs %d, total matvecs %d (lanczos steps %d) lam2 %s res %s (%.1fs)'%(m1,K,target,a1,d1,p2['steps'],p2['nmv'],tot,m1+p2['steps'],p2['lam2'],p2['res'],time.time()-t))
