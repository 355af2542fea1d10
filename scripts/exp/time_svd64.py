import sys, torch
sys.path.insert(0,'/root/repo')
import amp_sparc_spatialmodulation_b200 as pkg
F=8192
H=torch.view_as_complex(torch.randn(F,64,128,2,device='cuda')*0.09)
for _ in range(2): U,s,Vh=pkg.svd_batched(H)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); U,s,Vh=pkg.svd_batched(H); e1.record(); torch.cuda.synchronize()
print("svd 64x128 with U:", e0.elapsed_time(e1), "ms", F/e0.elapsed_time(e1)*1e3, "matrices/s", float((U@torch.diag_embed(s.to(torch.complex64))@Vh - H).abs().max()))
