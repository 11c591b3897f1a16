"""Bisect which part of the train step invalidates a CUDA-graph capture (run on a GPU box)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import algonauts2025_b200
from algonauts2025_b200 import ops
from algonauts2025_b200.model import FmriEncoder, FmriEncoderConfig
from algonauts2025_b200.pl_module import BrainModule
from algonauts2025_b200.segment import SegmentData, synthetic_batch
from algonauts2025_b200.trainer import MiniTrainer, default_optimizer

SMALL_DIMS = {"text": (2, 96), "audio": (2, 40), "video": (1, 72)}
torch.manual_seed(0)
model = FmriEncoder(SMALL_DIMS, 200, 25, FmriEncoderConfig(n_subjects=3, modality_dropout=0.0), hidden=384, depth=2, heads=6)
module = BrainModule(model=model, loss=torch.nn.MSELoss(), optim_config=None, metrics={}, max_epochs=1)
opt, sched = default_optimizer(model.parameters(), total_steps=100, model=model)
tr = MiniTrainer(module, opt, sched)
spec = tuple((k, v[0], v[1]) for k, v in SMALL_DIMS.items())
b = synthetic_batch(batch_size=3, t=74, t_out=25, n_outputs=200, n_subjects=3, seed=1, dims=spec)
dev = SegmentData(data={k: v.cuda() for k, v in b.data.items()}, segments=b.segments)
for _ in range(2):
    tr.eager_step(dev)
opt.init_all_state()
model.flush_subject_check()
torch.cuda.synchronize()


def stage(name, fn, mode="global"):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print("OK  ", name, mode, flush=True)
    except Exception as e:  # noqa: BLE001
        print("FAIL", name, mode, repr(e)[:300], flush=True)
        traceback.print_exc()
        torch.cuda.synchronize()


module.train()
def f_check():
    model._subjects(dev)
def f_fwd_nograd():
    with torch.no_grad():
        model(dev)
def f_fwd():
    model(dev)
def f_loss():
    module.training_step(dev, 0)
def f_bwd():
    opt.zero_grad(set_to_none=True)
    module.training_step(dev, 0).backward()
def f_full():
    opt.graph_begin()
    tr.run_step_body(dev)
    opt.graph_end()
for mode in ("global",):
    for name, fn in (("bwd", f_bwd), ("full", f_full)):
        stage(name, fn, mode)
