import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from fock_rs_b200.fock import FockBuilder
from fock_rs_b200 import _lib
from helpers import flat_for, random_symmetric_density
for mol,basis in [('benzene','6-31G'),('w16','6-31G')]:
    fb=flat_for(mol,basis)
    D=random_symmetric_density(fb.n_basis)
    for screen in (0,1):
        t=time.time()
        with FockBuilder(fb,screen=screen) as b:
            t1=time.time(); b.eri_build(); t2=time.time()
            b.eri_build(); t3=time.time()
            G=b.build_g(D); G=b.build_g(D); t4=time.time()
            for _ in range(5): G=b.build_g(D)
            print(mol,basis,'screen',screen,'create %.3f build1 %.3f build2 %.3f'%(t1-t,t2-t1,t3-t2), b.timings(), b.stats())
