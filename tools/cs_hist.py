import os, sys, torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
S = 512
seg = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=(32, 32, 16), seed=0, dtype=torch.int32, order="F")
os.environ["SYK_CS_DEBUG"] = "1"
out = dev.detect_cs(seg)
torch.cuda.synchronize()
