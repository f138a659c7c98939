import os, sys
sys.path.insert(0, "lrp-imagecaptioning-pytorch_b200")
import torch, torch.nn.functional as F
from lrpx import tc
torch.backends.cudnn.allow_tf32 = False
def run(n,h,w,cin,cout):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n,cin,h,w,generator=g).to(torch.bfloat16).float().cuda()
    wt = (torch.randn(cout,cin,3,3,generator=g)*0.1).to(torch.bfloat16).float().cuda()
    out = torch.zeros(tc.pf_rows(n,h,w), cout, device="cuda")
    tc.tc_conv(tc.nchw_to_pf(x), tc.weight_prep(wt,0), n,h,w,cin,cout,3,tc.EPI_STORE_F32,out)
    got = out.view(n,h+1,w+1,cout)[:,1:,1:,:].permute(0,3,1,2)
    ref = F.conv2d(x,wt,None,1,1)
    return float((got-ref).abs().max()/ref.abs().max())
for bo in ("1","0"):
    os.environ["LRPX_TC_SLAB"]="1"; os.environ["LRPX_TC_BASEOFF"]=bo
    for shp in [(1,8,8,64,64),(1,7,7,64,64),(2,14,14,64,512),(1,112,112,64,64),(1,224,224,64,64)]:
        try:
            print("baseoff",bo,shp,"relerr",run(*shp))
        except Exception as e:
            print("baseoff",bo,shp,"ERR",e); break
