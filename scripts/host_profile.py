"""cProfile of the HOST side of a training step (which Python calls the launch time goes to): `python scripts/host_profile.py dasm [batch]`."""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from transformer4sed_b200 import functional as F  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dasm"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"matsed": 64, "matsed_finetune2": 64, "pmam": 32, "dasm": 8}[name]
F.set_precision("bf16")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net, ext, arena, train_step, desc, wav_host = bench.build_workload(name, "bf16", B, dev, 0)
wav = wav_host.to(dev)
for _ in range(3):
    train_step(wav)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    train_step(wav)
    torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
