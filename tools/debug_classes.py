"""Debug aid: max |ERI(gpu) - ERI(oracle)| grouped by the angular momenta of the four functions."""
import sys, collections
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle
from fock_rs_b200.fock import FockBuilder
from helpers import flat_for
mol, basis = sys.argv[1], sys.argv[2]
screen = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fb = flat_for(mol, basis)
ob = oracle.OracleBasis.from_flat(fb)
want, _ = oracle.from_basis_packed(ob, oracle.FAITHFUL | (oracle.NO_CUT if screen == 1 else 0), fast=True)
with FockBuilder(fb, screen=screen) as b:
    b.eri_build()
    got = b.eri_export_packed()
n = fb.n_basis
L = np.asarray(fb.angular).reshape(n, 3).sum(1)
err = collections.defaultdict(float); cnt = collections.Counter(); worst = {}
tri = lambda a, b: a * (a + 1) // 2 + b
for x in range(n):
    for y in range(x + 1):
        xy = tri(x, y)
        for z in range(x + 1):
            for w in range(z + 1):
                zw = tri(z, w)
                if zw > xy: continue
                e = abs(got[tri(xy, zw)] - want[tri(xy, zw)])
                a, bb = sorted((L[x], L[y]), reverse=True); c, d = sorted((L[z], L[w]), reverse=True)
                key = max((a, bb), (c, d)) + min((a, bb), (c, d))
                cnt[key] += 1
                if e > err[key]:
                    err[key] = e; worst[key] = (x, y, z, w, got[tri(xy, zw)], want[tri(xy, zw)])
for k in sorted(cnt):
    print(k, cnt[k], "max err %.3e" % err[k], worst.get(k))
