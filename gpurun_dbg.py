import sys; sys.path.insert(0,'tests')
import numpy as np, oracle, cvr_b200
from cvr_b200 import gen
from helpers import to_oracle_csr
d = gen.random_sparse(20000, 15000, 150000, seed=41, empty_frac=0.25)
csr = to_oracle_csr(d)
x = np.random.default_rng(7).uniform(-1, 1, csr.n_cols + 1)
yc, mag = oracle.csr_spmv(csr, x)
for T in [1, 7, 64, 1000, 5000]:
    with cvr_b200.CvrMatrix(d.to_host(), T) as m:
        for rep in range(3):
            y,_ = m.spmv(x)
            err = np.abs(y-yc); bad = np.flatnonzero(err > 1e-12*mag + 1e-300)
            print('T',T,'rep',rep,'bad rows',bad.size, bad[:8], 'len/chunk', csr.nnz//T)
