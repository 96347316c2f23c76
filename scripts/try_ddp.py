import os, sys, time, torch, torch.distributed as dist
ROOT="/root/repo"; sys.path[:0]=[ROOT, ROOT+"/mamba-unet_b200"]
from selscan_b200 import workloads as wl
from selscan_b200.vssm import DiceLoss, MambaUnet
local=int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1337)
x=torch.rand(24,1,224,224,device=dev); y=torch.randint(0,4,(24,224,224),device=dev)
dice=DiceLoss(4)
def run(tag, **kw):
    model=MambaUnet(num_classes=4).to(dev).train()
    net=torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], **kw) if kw is not None else model
    opt=wl.make_sgd(net)
    for _ in range(5): wl.supervised_step(net,opt,dice,x,y)
    torch.cuda.synchronize(); dist.barrier()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(15): wl.supervised_step(net,opt,dice,x,y)
    e1.record(); torch.cuda.synchronize()
    if dist.get_rank()==0: print(tag, round(e0.elapsed_time(e1)/15,3),"ms", flush=True)
    del net,opt,model
run("nodist", **{}) if False else None
run("default")
run("bucket_view", gradient_as_bucket_view=True)
run("static", gradient_as_bucket_view=True, static_graph=True)
run("cap100", gradient_as_bucket_view=True, bucket_cap_mb=100)
run("cap10", gradient_as_bucket_view=True, bucket_cap_mb=10)
dist.destroy_process_group()
