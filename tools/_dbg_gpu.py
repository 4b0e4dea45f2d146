import sys; sys.path.insert(0,'.')
import numpy as np, torch, warnings; warnings.filterwarnings('ignore')
from oracle import boxqp as bq
from model_predictive_control_b200 import problem
for N in (5,30):
    rng=np.random.default_rng(100+N); batch=200
    x0=np.stack([rng.uniform(-100,0,batch), rng.uniform(-10,15,batch)],1); x0[0]=[-100,0]; x0[1]=[-1,14]
    prob=problem.Problem(N=N); res=problem.LinearMPC(prob).solve(x0)
    st=res.status.cpu().numpy(); it=res.iters.cpu().numpy()
    op=bq.Problem(N=N); ulo,uhi,xlo,xhi=bq.problem_bounds(op)
    port=bq.ipm_riccati(op.A,op.B,op.Q,op.R,op.Q,N,x0,ulo,uhi,xlo,xhi)
    print(N,'gpu status',np.bincount(st,minlength=4),'port',np.bincount(port['status'],minlength=4))
    for b in np.nonzero((st!=port['status'])|(np.abs(it-port['iters'])>0))[0]:
        print(' b',b,x0[b],'gpu',st[b],it[b],'port',port['status'][b],port['iters'][b])
