import sys
sys.path.insert(0,'/root/repo')
import torch
from fastc_b200 import ECompressionFormat as F, lib
from fastc_b200.synth import synth_rgba_torch
g=lib(); q=int(sys.argv[1]); size=int(sys.argv[2])
img=synth_rgba_torch(size,size,1,opaque=True); out=torch.zeros((size//4)**2*8,dtype=torch.uint8,device='cuda')
g.compress_device(F.ETC1,img,out,width=size,height=size,etc1_quality=q); torch.cuda.synchronize()
