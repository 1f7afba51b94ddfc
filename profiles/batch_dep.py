import sys; sys.path[:0]=[".","tests","oracle"]
import numpy as np, helpers as H, squander_b200 as sq
n,L=10,4
c=H.adaptive_circuit(n,L); P=c.get_Parameter_Num()
U=np.ascontiguousarray(H.random_unitary(1<<n).conj().T); p=H.random_params(P,seed=1,batch=1)
e=sq.Engine(0); e.upload_matrix(U); e.set_circuit(c); e.set_cost(0,0)
ref=None
for B in (1,2,3,5,6,7,8,16,32):
    f,g=e.cost_grad_batched(np.repeat(p,B,axis=0))
    if ref is None: ref=g[0]
    print(B, "same as B=1:", (g[0]==ref).all(), "max diff", abs(g[0]-ref).max(), "within batch same:", all((g[i]==g[0]).all() for i in range(B)))
