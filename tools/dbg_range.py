import sys
import torch
sys.path.insert(0, ".")
import mininf_b200 as mininf
from mininf_b200.engine import abi
from oracle import configs
DEV = "cuda:0"
torch.manual_seed(1)
n, p, S = 4096, 64, 8
config = configs.regression(n, p, device=DEV, gen_device="cpu")
approx, _ = config.approximation(device=DEV)
module = mininf.nn.EvidenceLowerBoundLoss(S, check="off")
model = lambda: config.model(mininf)
print(float(module(mininf.condition(model, **config.data), approx)), module.last_plan.dense_sites[0][1], int(module.last_plan.status))
plan0 = module.last_plan
other = {k: v.clone() for k, v in config.data.items()}
other["X"][17, 3] = 1.0e5
print(float(module(mininf.condition(model, **other), approx)), module.last_plan is plan0, module.last_plan.dense_sites[0][1], int(module.last_plan.status), plan0.rebindable)
