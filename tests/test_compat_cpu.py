"""Boundaries B2 / B4 on the CPU: the compatibility environment makes the reference's own train.py importable on this
PyTorch, and the B200 loaders shard / split / batch exactly like the reference's `dataset_selector` (datasets.py:445-582)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="oracle/_ref not staged (oracle/make_ref.py)")


def _index(tmp_path, n):
    p = tmp_path / "index.csv"
    p.write_text("Filepath,Label\n" + "".join(f"f{i}.JPEG,{i % 1000}\n" for i in range(n)))
    return str(p)


def test_yacs_stand_in():
    sys.path.append(os.path.join(ROOT, "rgb_no_more_b200", "compat", "shims"))
    from yacs.config import CfgNode as CN
    c = CN()
    c.A = CN()
    c.A.B = 3
    d = c.clone()
    d.A.B = 4
    assert c.A.B == 3 and d.A.B == 4 and isinstance(d.A, CN)
    d.freeze()
    with pytest.raises(AttributeError):
        d.A.B = 5


def test_loader_surface_and_sharding(tmp_path):
    from rgb_no_more_b200 import ddp
    from rgb_no_more_b200.compat import loader as L
    idx = _index(tmp_path, 103)
    tr = [L.dataset_selector("imagenet_dct", "train", idx, "", batch_size=8, shuffle=True, distributed=True, rank=r, world_size=4)
          for r in range(4)]
    assert all(len(t) == 4 for t in tr)                               # ceil(ceil(103 / 4) / 8)
    for t in tr:
        t.sampler.set_epoch(3)
    got = [t._indices() for t in tr]
    assert got == [ddp.shard_indices(103, r, 4, train=True, epoch=3, seed=0) for r in range(4)]
    assert sorted(set(sum(got, []))) == list(range(103))             # padded cover, like DistributedSampler
    te = [L.dataset_selector("imagenet_dct", "test", idx, "", batch_size=8, distributed=True, rank=r, world_size=4) for r in range(4)]
    assert sorted(sum((t._indices() for t in te), [])) == list(range(103))      # DistributedEvalSampler: no image twice
    with pytest.raises(NotImplementedError):
        L.dataset_selector("imagenet", "train", idx)


@needs_ref
def test_split_and_samplers_match_reference_dataset_selector(tmp_path):
    """Same train / minival / train-val split and the same per-epoch shard order as the reference's own loaders."""
    code = f"""
import sys, json, torch
sys.path.insert(0, {ROOT!r})
from rgb_no_more_b200.compat import env, loader as L
env.activate({REF!r})
import datasets as ds
idx = {_index(tmp_path, 257)!r}
out = {{}}
for rank in (0, 1):
    ref = ds.dataset_selector(dataset='imagenet_dct', type='train', indexpath=idx, basepath='', batch_size=16, num_workers=1, shuffle=True,
                              trainval_split=0.05, return_indices=True, distributed=True, rank=rank, world_size=2, seed=11997733,
                              ops_list=['Cutout'], num_ops=2, ops_magnitude=9)
    ours = L.dataset_selector(dataset='imagenet_dct', type='train', indexpath=idx, basepath='', batch_size=16, num_workers=1, shuffle=True,
                              trainval_split=0.05, return_indices=True, distributed=True, rank=rank, world_size=2, seed=11997733,
                              ops_list=['Cutout'], num_ops=2, ops_magnitude=9)
    assert list(ref[3]) == list(ours[3]) and list(ref[4]) == list(ours[4]), 'split differs'
    assert [len(x) for x in ref[:3]] == [len(x) for x in ours[:3]], ([len(x) for x in ref[:3]], [len(x) for x in ours[:3]])
    for ep in (0, 5):
        ref[0].sampler.set_epoch(ep); ours[0].sampler.set_epoch(ep)
        # reference order = positions in the train subset; ours = positions in the same train-row list
        assert list(iter(ref[0].sampler)) == ours[0]._indices(), 'train shard order differs'
    assert list(iter(ref[1].sampler)) == ours[1]._indices() and list(iter(ref[2].sampler)) == ours[2]._indices()
    assert torch.initial_seed() == 11997733 + rank           # both leave the global generator seeded SEED + rank (datasets.py:523)
print('OK')
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@needs_ref
def test_reference_train_module_imports_and_parses(tmp_path):
    """train.py of the reference imports on this PyTorch through env.activate (yacs / timm / torchmetrics / torchinfo / fvcore /
    sysrsync stand-ins, dct_manip drop-in) and its scheduler construction (CosineAnnealingLR(verbose=...)) works."""
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from rgb_no_more_b200.compat import env, backend
env.activate({REF!r})
import train, dct_manip, torch
assert 'rgb_no_more_b200' in dct_manip.read_coefficients.__module__
sys.argv = ['train.py', '--domain', 'dct', '--embed_type', '1', '--model_arch', 'vits', '--train']
a = train.parse_args()
cfg = train.configs.generate_config(modelarch=a.model_arch, domain=a.domain, modelver=a.embed_type, subblock=True)
p = torch.nn.Parameter(torch.zeros(2, 2))
class M(torch.nn.Module):
    def __init__(s):
        super().__init__(); s.w = torch.nn.Linear(2, 2)
crit, opt, wdec, sched, scaler = train.utils.get_optim_and_criterion(cfg, M(), range(10))
assert type(sched).__name__ == 'CosineAnnealingLR'
backend.install(train.utils)
assert train.utils.get_model is backend.get_model and train.utils.unpack_data is backend.unpack_data
print('OK')
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]
