import sys; sys.path[:0]=[".","tests","oracle"]
import numpy as np, helpers as H, squander_b200 as sq
n,L=7,1
c=H.adaptive_circuit(n,L); P=c.get_Parameter_Num()
U=np.ascontiguousarray(H.random_unitary(1<<n).conj().T)[:, :16].copy(); p=H.random_params(P,seed=1,batch=2)
e=sq.Engine(0); e.upload_matrix(U); e.set_circuit(c); e.set_cost(0,0)
f1,g1=e.cost_grad_batched(p)
print(f1, e.last_kernel_time())
