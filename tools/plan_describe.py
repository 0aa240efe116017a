import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, qcpinn_b200 as qb
F=qb.functional
for a,n,L,dt in [("sim_circ_15",16,2,torch.float32),("cross_mesh",10,2,torch.float32),("cascade",12,1,torch.float32),("cross_mesh",13,1,torch.float32),("sim_circ_15",16,2,torch.float64),("layered",8,2,torch.float32)]:
    prog=qb.program.compile_program(a,n,L,None)
    print(a, len(prog.ops), F.Plan(prog,0,dt,50,torch.device("cuda",0)).describe())
