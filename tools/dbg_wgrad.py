import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
pkg = load_package()
B,h,w,C,F,k = 1,8,32,32,32,3
layer = pkg.conv2d(F, kernel_size=k)
x = torch.randn(B,h,w,C,device='cuda'); dy = torch.randn(B,h,w,F,device='cuda')
layer.build(tuple(x.shape))
dx, dk, db = pkg.distortion_aware_ops.conv2d_backward(layer, x, dy)
torch.cuda.synchronize()
print('dk abs sum', dk.abs().sum().item(), 'nonzero', (dk!=0).sum().item(), 'of', dk.numel())
print('dx abs sum', dx.abs().sum().item())
