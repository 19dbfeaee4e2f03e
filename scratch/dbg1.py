import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, oracle
from fock_rs_b200.fock import FockBuilder
from helpers import flat_for
fb=flat_for('water_dimer','STO-3G'); n=fb.n_basis
ob=oracle.OracleBasis.from_flat(fb)
want,_=oracle.from_basis_packed(ob,0,True)
with FockBuilder(fb,kernel=0) as b: got=b.eri_build().eri_export_packed(); print(b.stats())
bad=np.nonzero(np.abs(got-want)>1e-10)[0]
print(len(bad))
def untri(k):
    r=int((np.sqrt(8*k+1)-1)/2)
    while r*(r+1)//2>k: r-=1
    while (r+1)*(r+2)//2<=k: r+=1
    return r,k-r*(r+1)//2
ang=fb.angular.sum(1)
from collections import Counter
cnt=Counter()
for e in bad[:4000]:
    xy,zw=untri(e); x,y=untri(xy); z,w=untri(zw)
    cnt[(ang[x],ang[y],ang[z],ang[w], got[e]==0, want[e]==0)]+=1
for k,v in sorted(cnt.items()): print(k,v)
for e in bad[:10]:
    xy,zw=untri(e); x,y=untri(xy); z,w=untri(zw)
    print(x,y,z,w,got[e],want[e])
