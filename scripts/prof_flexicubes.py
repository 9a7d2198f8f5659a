"""One FlexiCubes get_geometry forward + backward at R = 128 for an ncu capture of the fc_* kernels:
ncu --set full --clock-control none --import-source on -k regex:fc_ -o gpurun_out/prof_r1h_flexicubes python scripts/prof_flexicubes.py"""
import sys

sys.path.insert(0, ".")
import torch

from geosplatting_b200.flexicubes import FlexiCubes

dev, R = "cuda:0", 128
fc0 = FlexiCubes.from_resolution(R, random_sdf=False, scale=0.9, device=dev)
gv = fc0.vertices
sdf = (gv.norm(dim=-1, keepdim=True) - 0.6 + 0.06 * torch.sin(5.0 * gv[:, :1]) * torch.cos(4.0 * gv[:, 1:2]))
sdf = sdf.clone().requires_grad_(True)
w = (0.1 * torch.randn(fc0.indices.shape[0], 21, device=dev)).requires_grad_(True)
fc = fc0.replace(sdf_values=sdf, alpha=w[:, :8], beta=w[:, 8:20], gamma=w[:, 20:])
mesh, l_dev = fc.dual_marching_cubes()
loss = mesh.vertices.square().sum() + l_dev.mean() + fc.compute_entropy()
torch.autograd.grad(loss, [sdf, w])
torch.cuda.synchronize()
print("faces", mesh.indices.shape[0])
