import sys, os, json
sys.path[:0]=['/root/repo','/root/repo/dnn-compression-tensor-admm_b200','/root/repo/scripts']
import torch, workloads, hp_tables
from admm import ADMM
import bench_configs as bc
for key in ('resnet32_tk','resnet32_tk2','resnet32_tt'):
    wb,hb,fmt=workloads.CONFIGS[key]
    model=workloads.ParamBag(wb(seed=0), device='cuda:0'); hp=hb()
    admm=ADMM(model,1e-3,hp.fresh() if hasattr(hp,'fresh') else hp,fmt,'cuda:0')
    print(key, round(bc.time_updates(admm,5),3), {k:os.environ.get(k) for k in ('TTA_GRAM_TC','TTA_EIG_SOLVER','TTA_GRAM_IN_PLACE')})
