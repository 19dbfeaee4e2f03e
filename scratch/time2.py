import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from fock_rs_b200.fock import FockBuilder
from fock_rs_b200 import _lib
from helpers import flat_for, random_symmetric_density
for mol,basis,screen in [('alkane40','6-31G',0),('alkane40','6-31G',1)]:
    fb=flat_for(mol,basis)
    D=random_symmetric_density(fb.n_basis)
    with FockBuilder(fb,screen=screen,storage=_lib.STORAGE_DIRECT) as b:
        t0=time.time(); b.eri_build(); t1=time.time()
        G=b.build_g(D); t2=time.time()
        G=b.build_g(D); t3=time.time()
        print(mol,basis,'screen',screen,'build %.3f g1 %.3f g2 %.3f'%(t1-t0,t2-t1,t3-t2), b.timings(), b.stats(), flush=True)
