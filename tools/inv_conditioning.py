"""CPU experiment behind the explicit-inverse choices of the fused step kernel: error of the posterior
variance (relative to outputscale, in units of the 1e-9 parity tolerance) for substitution, 8/16/24-row block inverses
and the full inverse of L_oo, against an 80-bit reference.   python tools/inv_conditioning.py"""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
import torch
from sampling_gpmpc_b200 import configs
from sampling_gpmpc_b200.envs import make_env_spec
from sampling_gpmpc_b200.agent import gp_hypers_from_params
np.random.seed(0)
def run(params, T_use):
    spec=make_env_spec(params); X,Y=spec.initial_training_data(params)
    X=X.numpy(); d=X.shape[1]
    ls,os_,noise=gp_hypers_from_params(params, spec.g_ny, d, use_grad=True)
    obs=~np.isnan(Y.numpy()[0])  # (n,T)
    pts,tasks=np.nonzero(obs)
    for j in range(spec.g_ny):
        l=ls[j]; o=os_[j]
        def cov(xa,ta,xb,tb):
            r=xa-xb; k=o*np.exp(-0.5*np.sum((r/l)**2))
            if ta==0 and tb==0: return k
            if ta==0: return k*r[tb-1]/l[tb-1]**2
            if tb==0: return -k*r[ta-1]/l[ta-1]**2
            h=-(r[ta-1]/l[ta-1]**2)*(r[tb-1]/l[tb-1]**2)
            if ta==tb: h+=1/l[ta-1]**2
            return k*h
        m=len(pts)
        A=np.array([[cov(X[pts[a]],tasks[a],X[pts[b]],tasks[b]) for b in range(m)] for a in range(m)])
        A+=np.diag(noise[j][tasks])
        L=np.linalg.cholesky(A)
        Lq=L.astype(np.longdouble)
        lo,hi=X.min(0),X.max(0)
        errs={}
        for trial in range(200):
            xs=lo+(hi-lo)*np.random.rand(d)
            K=np.array([[cov(X[pts[a]],tasks[a],xs,tb) for tb in range(T_use)] for a in range(m)])
            # truth in longdouble substitution
            Kq=K.astype(np.longdouble); wq=np.zeros_like(Kq)
            for i in range(m): wq[i]=(Kq[i]-Lq[i,:i]@wq[:i])/Lq[i,i]
            def sub(): 
                w=np.zeros_like(K)
                for i in range(m): w[i]=(K[i]-L[i,:i]@w[:i])/L[i,i]
                return w
            def blockinv(bs):
                w=np.zeros_like(K)
                for i0 in range(0,m,bs):
                    i1=min(m,i0+bs)
                    rhs=K[i0:i1]-L[i0:i1,:i0]@w[:i0]
                    Dinv=np.linalg.inv(L[i0:i1,i0:i1])
                    w[i0:i1]=Dinv@rhs
                return w
            for name,w in (('sub',sub()),('b8',blockinv(8)),('b16',blockinv(16)),('b24',blockinv(24)),('full',blockinv(m))):
                S=(w.T@w); Sq=(wq.T@wq)
                e=np.max(np.abs(S-Sq).astype(np.float64))/o   # variance error relative to outputscale
                errs[name]=max(errs.get(name,0),e/1e-9)
        print(params['env']['dynamics'],'out',j,'m',m,'cond(L)=%.2e'%np.linalg.cond(L),{k:'%.2e'%v for k,v in errs.items()})
run(configs.car_residual_fs(8,5,True),3)
p=configs.pendulum2D_rollout(); run(p,4)
run(configs.pendulum1D_sqp(),3)
