"""Boundary B4 on the GPU: the REFERENCE's train.py (CLI, config, DDP wrap, loop, evaluation, checkpoint, final save) driven by
`python -m rgb_no_more_b200.compat.launch --backend b200` on synthetic 512x512 JPEG files: `--domain dct --model_arch vits`,
one epoch of 3 optimiser steps + minival / train-val / test evaluation, B200 loaders + B200 model behind the reference's seams."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="oracle/_ref not staged (oracle/make_ref.py)")


def _dataset(tmp_path, rows=120):
    from rgb_no_more_b200 import synth
    data = tmp_path / "data"
    data.mkdir()
    n_files = 16
    for i in range(n_files):
        (data / f"img_{i}.JPEG").write_bytes(synth.synth_jpeg(i))
    for name in ("train", "val"):
        with open(tmp_path / f"index_{name}.csv", "w") as f:
            f.write("Filepath,Label\n")
            for r in range(rows if name == "train" else 40):
                f.write(f"img_{r % n_files}.JPEG,{r % 8}\n")
    return data


def _launch(tmp_path, data, modes, arch, batch, extra, save, gpus=1):
    cmd = [sys.executable, "-m", "rgb_no_more_b200.compat.launch", "--ref", REF, "--backend", "b200", "--",
           *modes, "--domain", "dct", *([] if "--embed_type" in extra else ["--embed_type", "1"]), "--model_arch", arch,
           "--batch", str(batch), "--epochs", "1",
           "--warmup_steps", "2", "--num_gpus", str(gpus), "--num_cpus", "4", "--no_extract", "--no_resize", "--temp_datapath", str(data),
           "--indexpaths", f"{tmp_path / 'index_train.csv'},{tmp_path / 'index_val.csv'}", "--savepath", str(save), "--verbose", "1",
           "--port", str(_free_port()), "--num_ops", "2", "--ops_magnitude", "9"] + extra
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), RGBNM_CHILD_LOG="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=env, cwd=str(tmp_path))
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-4000:]
    return log


def _test_line(log):
    lines = [ln for ln in log.splitlines() if "[Test Acc:" in ln]
    assert lines, log[-3000:]
    return lines[-1][lines[-1].index("[Test Acc:"):]


@needs_ref
@pytest.mark.parametrize("arch,batch,extra", [("vits", 40, []), ("vits", 40, ["--amp", "1", "--ampdtype", "bf16"]), ("swinv2", 20, []),
                                              ("vits", 40, ["--benchmark", "2"]), ("vits", 40, ["--embed_type", "2"]),
                                              ("vits", 40, ["--embed_type", "2", "--no_subblock"]), ("vitti", 40, ["--embed_type", "1", "--no_subblock"])])
def test_reference_train_py_runs_on_the_b200_backend(tmp_path, arch, batch, extra):
    """ViT-S with the reference's defaults (AMP off) and under its autocast + GradScaler loop (`--amp 1 --ampdtype bf16`), SwinV2-T
    (AMP on by default, utils/configs.py:137): the engines ignore the caller's autocast region.  `--benchmark 2` additionally runs
    the reference's own loader / model / pipeline throughput loops (eval.py:53-180, benchmark.py) over the B200 loaders and model."""
    data = _dataset(tmp_path)
    save = tmp_path / "out" / "model.pth"
    save.parent.mkdir()
    log = _launch(tmp_path, data, ["--train", "--eval"], arch, batch, extra, save)
    banner = f"rgbnm B200 backend: ViT ({arch})" if arch != "swinv2" else "rgbnm B200 backend: SwinTransformerV2 (swinv2)"
    assert banner in log, log[-3000:]                                           # the opt-in took effect in the spawned rank
    _test_line(log)                                                             # the reference's evaluation ran to the end
    if "--benchmark" in extra:
        for line in ("Train loader:", "Model F/B pass:", "Train pipeline:", "Test pipeline:"):
            assert line in log, log[-3000:]                                     # eval.py:171-177: its FPS table, from the B200 path
    sd = torch.load(save, map_location="cpu")
    if arch != "swinv2":
        from rgb_no_more_b200 import vit as V
        emb, heads = (384, 6) if arch == "vits" else (192, 3)
        ver = int(extra[extra.index("--embed_type") + 1]) if "--embed_type" in extra else 1
        twin = V.ViT(patch_size=16, emb_size=emb, depth=12, n_classes=1000, drop_p=0.0, num_heads=heads, head_size=64, pixel_space="DCT",
                     ver=ver, use_subblock="--no_subblock" not in extra)
        assert sorted(sd.keys()) == sorted(twin.state_dict().keys())                 # = the reference's key set for this embedding
        assert sd["encoder.0.0.fn.eb_mha.qkv.weight"].shape == (3 * emb, emb)
    else:
        assert len(sd) == 250 and sd["layers.0.blocks.0.attn.qkv.weight"].shape == (288, 96)      # reference key set (swinv2.py)
    assert all(torch.isfinite(v.float()).all() for v in sd.values() if v is not None)
    ckdir = save.parent / "checkpoints"                                         # per-epoch checkpoint written by the reference loop
    assert ckdir.is_dir() and os.listdir(ckdir), os.listdir(save.parent)


@needs_ref
def test_eval_only_from_the_saved_checkpoint_reproduces_the_test_metrics(tmp_path):
    """`--eval --loadpath` (train.py:207-208, utils.load_model_and_report): the checkpoint the reference saved after training on the
    B200 backend, loaded into a fresh B200 model by the reference's own loader code, gives the same test accuracy / loss line."""
    data = _dataset(tmp_path)
    save = tmp_path / "out" / "model.pth"
    save.parent.mkdir()
    first = _test_line(_launch(tmp_path, data, ["--train", "--eval"], "vits", 40, [], save))
    again = _test_line(_launch(tmp_path, data, ["--eval"], "vits", 40, ["--loadpath", str(save)], tmp_path / "out" / "unused.pth"))
    assert first == again, (first, again)


@needs_ref
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_through_the_reference_launcher(tmp_path):
    """`--num_gpus 2`: the reference spawns two ranks, wraps the B200 model in DistributedDataParallel (train.py:137) and shards the
    index with its samplers' arithmetic; the run must finish and save finite weights.  (600 index rows: the reference's minival
    split is 1 % of the index, and its evaluate_model needs at least one batch on every rank, eval.py:49.)"""
    data = _dataset(tmp_path, rows=600)
    save = tmp_path / "out" / "model.pth"
    save.parent.mkdir()
    log = _launch(tmp_path, data, ["--train", "--eval"], "vits", 40, [], save, gpus=2)
    _test_line(log)
    sd = torch.load(save, map_location="cpu")
    assert len(sd) == 152 and all(torch.isfinite(v).all() for v in sd.values())
