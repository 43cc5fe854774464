import sys; sys.path.insert(0,'.')
import numpy as np, enoki_b200 as ek
x=np.linspace(0.9,1.1,9).astype(np.float32)
X=ek.Float32.copy(x)
r=ek.hprod(X)
print("hprod", r.numpy(), np.prod(x.astype(np.float64)))
X2=ek.Float32.copy(x)
r2=ek.hprod(X2); w = r2 / X2
print("w", w.numpy(), np.prod(x.astype(np.float64))/x)
X3=ek.Float32.copy(x)
z = ek.hprod(X3) + ek.hsum(X3)*2.0
print("z", z.numpy(), np.prod(x.astype(np.float64)) + 2*x.astype(np.float64).sum())
