import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from gopairingbasedcryptography_b200 import bn254
from oracle import bn254_ref as o, hash_to_curve_ref as h
e=bn254.default_engine()
for m in [b"", b"abc", b"a"*100]:
    g1=e.hash_to_g1_batch([m,m],h.DST_BYTES_G1)
    ref=o.g1_to_bytes(h.hash_to_g1(m,h.DST_BYTES_G1))
    p=o.g1_from_bytes(g1[0].tobytes())
    print(len(m),"g1 eq",g1[0].tobytes()==ref,"same both",(g1[0]==g1[1]).all(),"on curve",o.g1_on_curve(p) if p else None)
    g2=e.hash_to_g2_batch([m],h.DST_BYTES_G2)
    ref2=o.g2_to_bytes(h.hash_to_g2(m,h.DST_BYTES_G2))
    q=o.g2_from_bytes(g2[0].tobytes())
    print(len(m),"g2 eq",g2[0].tobytes()==ref2,"on curve",o.g2_on_curve(q) if q else None)
    # compare with neg
    if q: print("neg match", o.g2_to_bytes(o.g2_neg(q))==ref2)
    if p: print("g1 neg match", o.g1_to_bytes(o.g1_neg(p))==ref)
