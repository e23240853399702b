import os, sys, traceback
import torch
sys.path.insert(0, "/root/repo")
from edl_b200.models import ResNetVd, to_train_dtype
from edl_b200.trainer import StudentTrainer
dev = torch.device("cuda", 0)
torch.manual_seed(0)
def run(tag, same_model, first_graph):
    try:
        m = to_train_dtype(ResNetVd(18, class_dim=16, width_mult=0.25), torch.bfloat16, dev).train()
        x = torch.randn(8, 3, 32, 32).bfloat16().contiguous(memory_format=torch.channels_last).pin_memory()
        t = torch.softmax(torch.randn(8, 16), -1).bfloat16().pin_memory()
        for use_graph in (first_graph, True):
            if not same_model:
                m = to_train_dtype(ResNetVd(18, class_dim=16, width_mult=0.25), torch.bfloat16, dev).train()
            tr = StudentTrainer(m, 8, image_shape=(3, 32, 32), num_classes=16, lr=0.05, use_graph=use_graph, bucket_cap_mb=0.25)
            for _ in range(4):
                tr.step(x, t)
            torch.cuda.synchronize()
        print(tag, "OK", flush=True)
    except Exception as e:
        print(tag, "FAIL", str(e).splitlines()[0][:120], flush=True)
        torch.cuda.synchronize()
run("same model eager->graph", True, False)
run("fresh model eager->graph", False, False)
run("same model graph->graph", True, True)
