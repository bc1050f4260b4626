"""The unchanged reference script classes run on top of the drop-in overlay (construction + state-dict schema; CPU).
Needs the read-only reference checkout, which exists in the build container only -> skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import sys, types, torch
from lavender_b200.run import install_overlay
install_overlay("/root/reference")
# the reference's dataset.py / logger.py drag in packages that are not installed here (skimage, tensorboardX, ...):
class _DS(torch.utils.data.Dataset):
    def __init__(self, *a, **k): pass
m = types.ModuleType("dataset")
m.Dataset_Base, m.get_dl, m.get_tsv_dls, m.MetaLoader = _DS, (lambda *a, **k: None), None, object
from lavender_b200.agent import move_to_cuda
m.move_to_cuda = move_to_cuda
sys.modules["dataset"] = m
for name in ("tensorboardX", "deepspeed"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["tensorboardX"].SummaryWriter = object
import utils.lib as lib
assert lib.__file__.startswith(sys.argv[1]), lib.__file__            # the overlay shadows the hub
import model, agent, visbackbone.video_swin as vs
assert model.__file__.startswith(sys.argv[1]) and vs.__file__.startswith(sys.argv[1])
from main_pretrain_mlm import LAVENDER_Pretrain_MLM, Agent_Pretrain_MLM    # UNCHANGED reference script
assert sys.modules["main_pretrain_mlm"].__file__.startswith("/root/reference")
from lavender_b200.pretrain import FakeTokenizer, default_args
args = default_args(vis_backbone_size="tiny", size_batch=2, bert_config=None, dataset=["webvid2.5m"])
mdl = LAVENDER_Pretrain_MLM(args, FakeTokenizer())
import lavender_b200.model as M, lavender_b200.bert as B
assert isinstance(mdl, M.LAVENDER_Base) and isinstance(mdl.fc_mtm, B.BertOnlyMLMHead) and isinstance(mdl.trsfr, B.BertEncoder)
sys.path.insert(0, sys.argv[2])
import lavender_oracle as O
want = {k: tuple(s) for k, s in O.state_dict_schema(O.ModelCfg(swin=O.SWIN["tiny"], bert_layers=12))}
got = {k: tuple(v.shape) for k, v in mdl.state_dict().items()}
assert got == want, set(got) ^ set(want)
ag = Agent_Pretrain_MLM(args, mdl)
assert type(ag).__mro__[-2].__module__ == "lavender_b200.agent"
print("DROPIN_OK", len(got))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_unchanged_reference_script_builds_on_overlay():
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT
    r = subprocess.run([sys.executable, "-c", SCRIPT, os.path.join(ROOT, "dropin"), os.path.join(ROOT, "oracle")],
                       cwd=REF, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
