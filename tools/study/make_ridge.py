"""ridge volume (u32 r^2 at ridge points) of the bench phantom recipe, written raw for painter_study.
usage: python tools/study/make_ridge.py N inverse out.u32"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bonej2_b200.phantom import phantom_shapes
from oracle import lt_oracle
n, inverse, out = int(sys.argv[1]), bool(int(sys.argv[2])), sys.argv[3]
lt_oracle.build()
vol = lt_oracle.phantom_raster(phantom_shapes(n, n, n, 1), n, n, n)
dist, sq = lt_oracle.edt(vol, inverse=inverse)
rg = lt_oracle.ridge(dist)
np.where(rg > 0, sq, 0).astype(np.uint32).tofile(out)
print("bv/tv", float((vol != 0).mean()), "ridge points", int((rg > 0).sum()), "max r^2", int(sq.max()))
