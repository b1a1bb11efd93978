import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gopairingbasedcryptography_b200 import bn254, schemes
from oracle import bn254_ref as o, port
import common
eng = bn254.default_engine()
n, m = 2048, 100
g1, g2 = port.generators()
sb = common.scalar_bytes(common.scalars(4096, seed=1, edges=False))
Pn = eng.g1_mul_base_batch(g1, sb); Qn = eng.g2_mul_base_batch(g2, sb[:32 * 256])
big = np.tile(Pn, (n * m // 4096 + 1, 1))
cy = big[: n * m].reshape(n, m, 64)
dj, djp, d = Qn[:m], Qn[m:2 * m], Qn[-1]
c = Pn[:n]
ctil = eng.pair_batch(Pn[:64], np.tile(Qn[:1], (64, 1)))
ctil = np.tile(ctil, (n // 64, 1))
deltas = sb.reshape(-1, 32)[:m]
pol = schemes.bsw07_policy_lines(eng, dj, djp, d, deltas)
for _ in range(2):
    out = schemes.bsw07_decrypt_batch(eng, cy, cy, dj, djp, c, d, ctil, deltas, lines=pol, folded=True)
print(out.shape)
