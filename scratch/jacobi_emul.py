import numpy as np
f=np.float32
def rr_pair(n, rnd, idx):
    m=n-1
    if idx==0: return rnd, m
    return (rnd+idx)%m, (rnd-idx+m)%m
def inner(W, sweeps=12, eps=f(1.1920929e-7), thr=f(0)):
    n=W.shape[0]; Q=np.eye(n,dtype=f); W=W.copy()
    for sw in range(sweeps):
        rot_any=False
        for rr in range(n-1):
            rots=[]
            for t in range(n//2):
                p,q=rr_pair(n,rr,t)
                app,aqq,apq=W[p,p],W[q,q],W[p,q]
                lim=max(thr, eps*np.sqrt(abs(app)*abs(aqq)))
                if abs(apq)>lim:
                    theta=(aqq-app)/(f(2)*apq)
                    tt=(f(1) if theta>=0 else f(-1))/(abs(theta)+np.sqrt(f(1)+theta*theta))
                    c=f(1)/np.sqrt(f(1)+tt*tt); s=tt*c; tau=s/(f(1)+c)
                    rots.append((p,q,s,tau,app-tt*apq,aqq+tt*apq))
            if not rots: continue
            rot_any=True
            for (p,q,s,tau,dpp,dqq) in rots:
                wp=W[:,p].copy(); wq=W[:,q].copy()
                W[:,p]=wp-s*(wq+tau*wp); W[:,q]=wq+s*(wp-tau*wq)
                qp=Q[:,p].copy(); qq=Q[:,q].copy()
                Q[:,p]=qp-s*(qq+tau*qp); Q[:,q]=qq+s*(qp-tau*qq)
            for (p,q,s,tau,dpp,dqq) in rots:
                wp=W[p,:].copy(); wq=W[q,:].copy()
                W[p,:]=wp-s*(wq+tau*wp); W[q,:]=wq+s*(wp-tau*wq)
                W[p,p]=dpp; W[q,q]=dqq; W[p,q]=0; W[q,p]=0
        if not rot_any: break
    return W,Q,sw
rng=np.random.default_rng(0)
bias=[];tr=[]
for trial in range(20):
    B=rng.standard_normal((32,32)); G=(B@B.T/32).astype(f)
    W,Q,sw=inner(G)
    QtQ=(Q.astype(np.float64).T@Q.astype(np.float64))
    bias.append(np.mean(np.diag(QtQ)-1)); tr.append((np.trace(W.astype(np.float64))-np.trace(G.astype(np.float64)))/np.trace(G.astype(np.float64)))
print('inner sweeps',sw,'mean diag(QtQ)-1: %.3e +- %.3e'%(np.mean(bias),np.std(bias)), ' rel trace drift %.3e +- %.3e'%(np.mean(tr),np.std(tr)))
# near-diagonal case (late sweeps): small off-diagonals
bias=[];tr=[]
for trial in range(20):
    d=rng.uniform(0.5,5,32); E=rng.standard_normal((32,32))*1e-3; G=(np.diag(d)+E+E.T).astype(f)
    W,Q,sw=inner(G)
    QtQ=(Q.astype(np.float64).T@Q.astype(np.float64))
    bias.append(np.mean(np.diag(QtQ)-1)); tr.append((np.trace(W.astype(np.float64))-np.trace(G.astype(np.float64)))/np.trace(G.astype(np.float64)))
print('near-diag: sweeps',sw,'mean diag(QtQ)-1: %.3e +- %.3e'%(np.mean(bias),np.std(bias)), ' rel trace drift %.3e +- %.3e'%(np.mean(tr),np.std(tr)))
