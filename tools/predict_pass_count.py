#!/usr/bin/env python3
"""How predictable is the pass count of the canopy Newton loop (func:2744-2877) from what a column knows at the start of a
step?  (ANALYSIS TOOL, CPU only: runs the oracle step by step on 20 000 CONUS columns and fits predictors; results in
profiles/r02_tuning.md section 1b.)  A warp runs to the maximum of its 32 lanes, so columns with similar counts want to
share a warp: this measures what a per-step regrouping inside a 256-column CTA could save."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from noahmp_b200 import layout as L, synth
from noahmp_b200.tables import ParamTables
from oracle.oracle import Oracle
t=ParamTables.default()
n=20000; ns=24
w=synth.make_workload("conus1km", ncol=n, nsteps=ns)
f=synth.forcing_all(w)
o=Oracle(t,w.opts,fast="fast")
st=w.state.copy()
kw=dict(dt=w.dt, zsoil=synth.ZSOIL, yearlen=w.yearlen, zref=w.zref, co2_frac=w.co2_frac, o2_frac=w.o2_frac, foln=w.foln)
X=[];Y=[]
prev_it=np.zeros(n); prev_f=f[0]
for s in range(ns):
    d=np.zeros((L.NDIAG,n),np.float32)
    tv0=st[L.S["TV"]].copy(); tg0=st[L.S["TG"]].copy(); eah0=st[L.S["EAH"]].copy()
    o.run(st, np.ascontiguousarray(f[s:s+1]), w.static_f, w.static_i, julian0=w.julian0+s*w.dt/86400., diag=d, nthreads=8, **kw)
    it=(d[L.D["ITER"]].view(np.int32)&0xff).astype(float)
    F=f[s]
    feats=np.stack([F[0]-prev_f[0], np.abs(F[0]-prev_f[0]), F[5]-prev_f[5], F[5], np.hypot(F[2],F[3]), F[4]-prev_f[4], F[6]-prev_f[6], F[8], tv0-F[0], tg0-F[0], prev_it, F[7], d[L.D["FVEG"]], (w.static_i&0xff).astype(float)],1)
    if s>=2:
        X.append(feats[it>0]); Y.append(it[it>0])
    prev_it=it; prev_f=F
X=np.concatenate(X); Y=np.concatenate(Y)
print(X.shape, np.bincount(Y.astype(int))[6:])
names=["dT","|dT|","dSW","SW","wind","dQ","dLW","cosz","TV-Tair","TG-Tair","prev_it","prcp","fveg","lutyp"]
for i,nm in enumerate(names): print(nm, round(np.corrcoef(X[:,i],Y)[0,1],3), round(np.corrcoef(X[:,i],Y>8)[0,1],3))
from sklearn.ensemble import HistGradientBoostingRegressor
m=len(Y)//2
idx=np.random.default_rng(0).permutation(len(Y)); tr,te=idx[:m],idx[m:]
g=HistGradientBoostingRegressor(max_iter=200).fit(X[tr],Y[tr])
p=g.predict(X[te]); print("GBR corr", np.corrcoef(p,Y[te])[0,1])
# how much would sorting by prediction reduce warp-max sum? simulate: groups of 256 columns: random order vs sorted by p
yt=Y[te]; 
def cost(order):
    y=yt[order][:len(yt)//256*256].reshape(-1,8,32)
    return (y.max(2)-6).sum()/ (y-6).sum()*32 , y.max(2).mean()
rng=np.random.default_rng(1)
print("random: lanes-equivalent waste ratio, mean warp max", cost(rng.permutation(len(yt))))
# sort within blocks of 256
blocks=rng.permutation(len(yt))[:len(yt)//256*256].reshape(-1,256)
srt=np.concatenate([b[np.argsort(p[b])] for b in blocks])
print("sorted by prediction within 256:", cost(srt))
srt2=np.concatenate([b[np.argsort(yt[b])] for b in blocks])
print("sorted by truth within 256:", cost(srt2))
print("---- cheap predictors, sorted within blocks of 256: (waste ratio, mean warp max)")
Xt=X[te]
def ev(score, label):
    srt=np.concatenate([b[np.argsort(score[b])] for b in blocks]); print(label, cost(srt))
ev(-Xt[:,4], "wind only")
ev(-Xt[:,4]+0.002*Xt[:,3], "wind + SW")
from sklearn.linear_model import LinearRegression
for cols,label in (([4],"lin wind"),([4,3,0],"lin wind,SW,dT"),([4,3,0,8,9,12],"lin +TV-T,TG-T,fveg"),([0,2,3,4,5,6,7,8,9,12,13],"lin all but prev")):
    Z=np.column_stack([X[:,cols], 1/np.maximum(X[:,4:5],0.5), np.log(np.maximum(X[:,4:5],0.5))])
    lr=LinearRegression().fit(Z[tr],Y[tr]); pr=lr.predict(Z[te]); print(label, "corr", round(np.corrcoef(pr,Y[te])[0,1],3)); ev(pr,label)
for cols,label in (([4],"gbr wind"),([4,3],"gbr wind,SW"),([4,3,0],"gbr wind,SW,dT"),([4,3,0,12,13],"gbr wind,SW,dT,fveg,lutyp"),([4,3,0,8,9],"gbr wind,SW,dT,TV-T,TG-T")):
    g2=HistGradientBoostingRegressor(max_iter=150).fit(X[tr][:,cols],Y[tr]); pr=g2.predict(Xt[:,cols]); print(label,"corr",round(np.corrcoef(pr,Y[te])[0,1],3)); ev(pr,label)
# bucketed (16 buckets) instead of full sort, GBR full predictor
q=np.quantile(p,np.linspace(0,1,17)[1:-1]); bk=np.digitize(p,q)
ev(bk+1e-6*rng.random(len(bk)),"gbr all, 16 buckets")
q=np.quantile(p,np.linspace(0,1,9)[1:-1]); bk=np.digitize(p,q)
ev(bk+1e-6*rng.random(len(bk)),"gbr all, 8 buckets")
