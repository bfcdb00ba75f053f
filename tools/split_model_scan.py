"""CPU model scan of the split-integer scheme of the two X contractions inside FULL oracle fits: X~ with Sx digits, the
factor-side operands (A, Y) with Sa digits, groups k + l < G kept -- parity of the fit against the golden vectors.
usage: OMP_NUM_THREADS=1 python tools/split_model_scan.py Sx Sa G [golden names...]   (tiny matmuls: more BLAS threads only contend)
Finding (round 1): X~ at 5 digits with A / Y at 6 (20 plane products, 11 % fewer delivered bytes) is no closer to the reference
than 5 / 5 (W: 4e-12 .. 1.3e-10 against 9e-12 .. 1.7e-10; 6 / 6: 7e-14 .. 2e-13): the static truncation of X~ dominates."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), ROOT]
import numpy as np, time
import corex_oracle as oc
from conftest import load_golden
from test_split_scheme import pow2_above, split_digits
R=254
def planes(a, scale, S):
    return [p.astype(np.float64) for p in split_digits(a, scale, S, R)]
def recomb(groups):
    G=len(groups)
    acc=groups[G-1]
    for g in range(G-2,-1,-1): acc=acc/R+groups[g]
    return acc/(R*R)
class XModel:
    def __init__(self, Sx, Sa, G):
        self.Sx,self.Sa,self.G=Sx,Sa,G; self.cache={}
    def xplanes(self, xt):
        key=id(xt)
        if key not in self.cache:
            sx=pow2_above(np.abs(xt).max())
            self.cache={key:(sx, planes(xt, sx, self.Sx), xt)}
        return self.cache[key][:2]
    def k1(self, xt, a):   # Y = X A^T
        sx,dx=self.xplanes(xt)
        sa=np.array([pow2_above(np.abs(r).max()) for r in a])
        da=planes(a, sa[:,None], self.Sa)
        groups=[np.zeros((xt.shape[0], a.shape[0])) for _ in range(self.G)]
        for k in range(self.Sx):
            for l in range(self.Sa):
                if k+l<self.G: groups[k+l]+=dx[k]@da[l].T
        return recomb(groups)*sx*sa[None,:]
    def k2(self, xt, y):   # X^T Y  (n x m)
        sx,dx=self.xplanes(xt)
        sy=np.array([pow2_above(np.abs(c).max()) for c in y.T])
        dy=planes(y, sy[None,:], self.Sa)
        groups=[np.zeros((xt.shape[1], y.shape[1])) for _ in range(self.G)]
        for k in range(self.Sx):
            for l in range(self.Sa):
                if k+l<self.G: groups[k+l]+=dx[k].T@dy[l]
        return recomb(groups)*sx*sy[None,:]
def install(model):
    def project_sumsq(xt,a):
        y=model.k1(xt,a); return y, np.einsum('lj,lj->j',y,y)
    def sigma_times(xt,u,eps):
        y=model.k1(xt,u); d=model.k2(xt,y)
        return (1-eps**2)*d.T/xt.shape[0]+eps**2*u
    orig=oc.moments_ns
    def moments_ns(xt,w,eps,quick=False,yscale=1.):
        # same as oracle but X products through the model
        n_samples=xt.shape[0]
        y,s=project_sumsq(xt,w)
        m={}
        m["uj"]=(1-eps**2)*s/n_samples+eps**2*np.sum(w**2,axis=1)
        if quick and np.max(m["uj"])>=1.: return None
        d=model.k2(xt,y)
        rho=(1-eps**2)*d.T/n_samples+eps**2*w
        ry=w.dot(rho.T)
        m["Y_j^2"]=yscale**2/(1.-m["uj"])
        np.fill_diagonal(ry,1)
        inv=1./(1.-rho**2); rinv=rho*inv
        qij=np.dot(ry,rinv); si=np.sum(rho*rinv,axis=0)
        qs=np.einsum('ki,ki->i',rinv,qij-si*rho)
        m["rho"],m["ry"],m["invrho"],m["rhoinvrho"]=rho,ry,inv,rinv
        m["Qij"],m["Si"],m["Qi-Si^2"]=qij,si,qs
        m["TC"]=np.sum(np.log(1+si))-0.5*np.sum(np.log(1+qs))+0.5*np.sum(np.log(1-m["uj"]))
        if not quick:
            m["MI"]=-0.5*np.log1p(-rho**2)
            m['I(Y_j ; X)']=0.5*np.log(m["Y_j^2"])-0.5*np.log(yscale**2)
            m["TCs"]=m["MI"].sum(axis=1)-m['I(Y_j ; X)']
        return m
    oc.project_sumsq=project_sumsq; oc.sigma_times=sigma_times; oc.moments_ns=moments_ns
def rel(a,b): return np.abs(np.asarray(a)-np.asarray(b)).max()/np.abs(np.asarray(b)).max()
if __name__=='__main__':
    Sx,Sa,G=(int(v) for v in sys.argv[1:4])
    names=sys.argv[4:] or ["readme_demo_f64","big5_l0_f64","syn_400x300x10_f64","syn_60x400x8_f64","outliers_missing_f64","standard_missing_f64","adni_l1_f64","adni_l0_f64"]
    install(XModel(Sx,Sa,G))
    for name in names:
        z,kw,x=load_golden(name)
        mdl=oc.OracleCorex(work_dtype=np.float64, **kw)
        if name.startswith('readme_demo'): x=np.random.random((100,50))
        t0=time.time(); mdl.fit(x)
        n=min(len(mdl.history['TC']),len(z['history_TC']))
        print('%-28s Sx=%d Sa=%d G=%d iters %d/%d  ws %.2e  TC-traj %.2e  %.0fs'%(name,Sx,Sa,G,len(mdl.history['TC']),len(z['history_TC']), rel(mdl.ws,z['ws']) if mdl.ws.shape==z['ws'].shape else -1, rel(mdl.history['TC'][:n],z['history_TC'][:n]), time.time()-t0), flush=True)
