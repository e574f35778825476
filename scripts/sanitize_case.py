"""Small end-to-end case for compute-sanitizer: every class kernel (s,p,d), screening on, tensor + RI modes, 1e ints."""
import sys

import numpy as np

sys.path.insert(0, ".")
from quantumlab_jl_b200 import libqlb, workloads
from quantumlab_jl_b200.b200shells import B200Shells
from quantumlab_jl_b200.resolutionidentity import B200RI

libqlb.init(0)
geo, shells, nocc = workloads.build("h2o2_631gs")
dev = B200Shells(shells)
rng = np.random.default_rng(0)
A = rng.standard_normal((dev.nbf, dev.nbf)) * 0.2
P = A + A.T
J, K = dev.build_jk(P, tau=1e-10)
F = dev.fock(np.zeros_like(P), P)
S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
blk = dev.eri_block(0, 5, 3, 9)
geo1, sh1, _ = workloads.build("h2o_631gs")
d1 = B200Shells(sh1)
t = d1.eri_tensor()
ri = B200RI(d1, sh1)
Jr, Kr = ri.build_jk(np.eye(d1.nbf))
print("sanitize case ok", float(J.sum()), float(K.sum()), float(t.sum()), float(Kr.sum()))
