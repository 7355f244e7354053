import torch, sys
sys.path.insert(0,'/root/repo')
from fitpack_b200.batch import Engine
from fitpack_b200._lib import lib
e=Engine(torch.device('cuda',0))
for mode,name in ((0,'dmul+dadd Gop/s'),(1,'dfma G/s'),(2,'ddiv G/s'),(3,'dsqrt(+add) G/s')):
    v=lib().fpb_probe_fp64_gops(e._h, mode)
    print(name, round(v,1), 'per clk per SM @1.965GHz: %.2f'%(v/148/1.965))
