"""Run eager forwards of the C2 batch for ncu: warm-up outside the profiled range, then
cudaProfilerStart .. one forward .. cudaProfilerStop (use `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paths  # noqa: E402,F401

sys.path.insert(0, os.path.join(paths.REPO, "tests"))
from helpers import build_model, inputs_for  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 8
yaml = sys.argv[2] if len(sys.argv) > 2 else "coco/interformer_coco_w48_pure_en6.yaml"
cfg, model, sd = build_model(yaml)
model = model.cuda()
model.use_cuda_graph = False
persons = int(sys.argv[3]) if len(sys.argv) > 3 else 4
length = [persons] * images
hw = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (256, 192)
x, pm = inputs_for(length, *hw)
x, pm = x.cuda(), pm.cuda()
for _ in range(2):
    model(x, pm, length)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(x, pm, length)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
