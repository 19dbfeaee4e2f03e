import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from fock_rs_b200.fock import FockBuilder
from fock_rs_b200 import _lib
from helpers import flat_for, random_symmetric_density
mol,basis,screen,storage = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
fb=flat_for(mol,basis)
D=random_symmetric_density(fb.n_basis)
with FockBuilder(fb,screen=screen,storage=storage,keep_eri=False) as b:
    b.eri_build(); G=b.build_g(D)
    t0=time.time(); b.eri_build(); t1=time.time(); G=b.build_g(D); t2=time.time()
    tm=b.timings(); st=b.stats()
    print(os.environ.get('TAG',''), mol, 'screen',screen,'storage',storage,'build %.2f ms g %.2f ms | class %.3f g_dev %.3f sort %.2f plan %.2f tiles %d TF %.2f'%(1e3*(t1-t0),1e3*(t2-t1),tm['class_kernels_ms'],tm['g_device_ms'],tm['host_sort_ms'],tm['host_plan_ms'],st['warp_tiles'], st['model_flops']/(tm['class_kernels_ms']*1e-3)*1e-12), flush=True)
    print({k: st[k] for k in ("shells", "shell_quartets_unique", "shell_quartets_evaluated", "general_shell_quartets_evaluated", "prim_quartets_evaluated", "launches_build", "launches_g")})
