"""Host-side restatements checked against the reference's OWN code wherever the reference tree is mounted (the build
container; on the GPU box these tests skip) and against the third-party code it calls (torch's DistributedSampler):

  train_loop.ModelSaver          <-> pretrain_src/utils/save.py:23-74          (imported)
  train_loop.MetaLoader          <-> pretrain_src/data/loader.py:18-75         (imported; the draw replayed)
  parallel.shard_indices         <-> torch.utils.data.distributed.DistributedSampler (data/loader.py:148-150)
  nav.rel_pos_fts                <-> pretrain_src/data/common.py:142-160       (source; the module needs pynvml/networkx)
  nav.angle_fts, nav_synth.view_angles <-> map_nav_src/utils/data.py:176-200   (source; the module needs jsonlines, spacy ...)
  nav.language_inputs            <-> map_nav_src/r2r/agent.py:63-90 `_language_variable` (source)
"""
import importlib.util
import math
import os
import re
import textwrap
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

import magic_b200  # noqa: F401
from magic_b200 import nav, nav_synth, parallel
from magic_b200.train_loop import MetaLoader, ModelSaver

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _source_functions(path, names, ns, method=False):
    src = open(path).read()
    ind = "    " if method else ""
    for name in names:
        m = re.search(rf"^{ind}def {name}\(.*?(?=^{ind}def |^{ind}class |\Z)", src, re.S | re.M)
        exec(textwrap.dedent(m.group(0)), ns)
    return ns


class _Opt:
    def state_dict(self):
        return {"step": 3, "x": torch.arange(4.0)}


@needs_ref
def test_model_saver_writes_what_the_reference_saver_writes(tmp_path):
    ref = _load(f"{REF}/pretrain_src/utils/save.py", "ref_save")
    model = nn.Sequential(nn.Linear(3, 2), nn.LayerNorm(2))
    wrapped = nn.Module()
    wrapped.module = model  # a DDP-like wrapper: the keys gain the `module.` prefix both savers strip
    a, b = tmp_path / "ref", tmp_path / "ours"
    a.mkdir(), b.mkdir()
    for saver in (ref.ModelSaver(str(a)), ModelSaver(str(b))):
        saver.save(wrapped, 7, _Opt())
        saver.save(model, 11)                       # no optimizer: no train_state file
        saver.save_latest(wrapped, 8, _Opt())
        saver.save_latest(wrapped, 9, _Opt(), is_max=True)
    assert sorted(os.listdir(a)) == sorted(os.listdir(b))
    for f in os.listdir(a):
        x, y = torch.load(a / f), torch.load(b / f)
        if f.startswith("train_state"):
            assert x["step"] == y["step"] and torch.equal(x["optimizer"]["x"], y["optimizer"]["x"])
        else:
            assert list(x) == list(y) and all(torch.equal(x[k], y[k]) and y[k].device.type == "cpu" for k in x)


@needs_ref
@pytest.mark.parametrize("accum", [1, 3])
def test_meta_loader_replays_through_the_reference_class(accum, monkeypatch):
    """Same (task, batch) sequence, same epoch restarts and `pre_epoch` calls as the reference MetaLoader when its
    per-step multinomial draw is replaced by our pre-agreed schedule (the one deliberate difference)."""
    ref = _load(f"{REF}/pretrain_src/data/loader.py", "ref_loader")
    calls_ref, calls_ours = [], []

    def loaders(log):
        return {"mlm": ([1, 2, 3], 2, lambda e: log.append(("mlm", e))), "sap": ([10], 1, lambda e: log.append(("sap", e))),
                "cfp": ([100, 200], 1, lambda e: log.append(("cfp", e)))}

    n = 40
    ours = MetaLoader(loaders(calls_ours), seed=9, num_steps=n, accum_steps=accum)
    seq = [x for x in ours]
    assert len(seq) == n and {t for t, _ in seq} == {"mlm", "sap", "cfp"}
    draws = iter([ours.tasks.index(t) for t in ours.schedule])
    monkeypatch.setattr(ref.torch, "multinomial", lambda p, k: torch.tensor([next(draws)]))
    r = ref.MetaLoader(loaders(calls_ref), accum_steps=accum, distributed=False, device="cpu")
    assert r.names == ours.tasks and r.sampling_ratios.tolist() == [2.0, 1.0, 1.0]
    it = iter(r)
    assert [next(it) for _ in range(n)] == seq
    assert calls_ref == calls_ours and len(calls_ours) > 3


@pytest.mark.parametrize("n,world,seed,drop_last", [(101, 4, 3, False), (101, 4, 3, True), (64, 8, 0, False),
                                                     (7, 2, 11, False), (1000, 3, 5, False)])
def test_shard_indices_are_distributed_sampler_indices(n, world, seed, drop_last):
    from torch.utils.data.distributed import DistributedSampler
    data = list(range(n))
    for shuffle in (True, False):
        got = [parallel.shard_indices(n, r, world, seed=seed, shuffle=shuffle, drop_last=drop_last) for r in range(world)]
        want = [list(DistributedSampler(data, num_replicas=world, rank=r, shuffle=shuffle, seed=seed, drop_last=drop_last))
                for r in range(world)]
        assert got == want
        if not drop_last:
            assert set(sum(got, [])) == set(data)


@needs_ref
def test_nav_geometry_follows_the_reference_functions():
    ns = _source_functions(f"{REF}/pretrain_src/data/common.py", ["calculate_vp_rel_pos_fts"], dict(np=np, math=math))
    rng = np.random.RandomState(3)
    for _ in range(200):
        a, b = rng.uniform(-10, 10, 3), rng.uniform(-10, 10, 3)
        bh, be = rng.uniform(-3, 3), rng.uniform(-1, 1)
        assert np.allclose(nav.rel_pos_fts(a, b, bh, be), ns["calculate_vp_rel_pos_fts"](a, b, bh, be), rtol=1e-12, atol=1e-12)
    assert np.allclose(nav.rel_pos_fts(a, a), ns["calculate_vp_rel_pos_fts"](a, a))  # coincident points: the 1e-8 floor
    ns = _source_functions(f"{REF}/map_nav_src/utils/data.py", ["get_angle_fts", "get_view_rel_angles"], dict(np=np, math=math))
    h, e = rng.uniform(-3, 3, 9), rng.uniform(-1, 1, 9)
    for size in (4, 8, 128):
        assert np.array_equal(nav.angle_fts(h, e, size), ns["get_angle_fts"](h, e, size))
    ref_angles = ns["get_view_rel_angles"](12)  # base view 12 = heading 0 on the horizontal row (agent.py:1409,1416)
    assert ref_angles.shape == (36, 2) and np.allclose(nav_synth.view_angles(), ref_angles, atol=1e-6)


class _NP:
    """numpy with the `np.bool` alias the reference source still uses (removed in numpy 1.24)."""
    bool = bool

    def __getattr__(self, k):
        return getattr(np, k)


@needs_ref
def test_language_inputs_match_the_reference_collator(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ns = _source_functions(f"{REF}/map_nav_src/r2r/agent.py", ["_language_variable"], dict(np=_NP(), torch=torch), method=True)
    rng = np.random.RandomState(1)
    obs = [{"instr_encoding": [0] + rng.randint(3, 50000, size=n).tolist() + [2]} for n in (5, 17, 1, 9)]
    want = ns["_language_variable"](types.SimpleNamespace(), obs, None, None)
    got = nav.language_inputs(obs, "cpu")
    assert torch.equal(got["txt_ids"], want["txt_ids"]) and got["txt_ids"].dtype == want["txt_ids"].dtype
    assert torch.equal(got["txt_masks"], want["txt_masks"])
    assert got["txt_lens"].tolist() == [len(o["instr_encoding"]) for o in obs]


def _agent_snippet(first, last):
    """Lines of the reference rollout (inline code, not a function) from the one containing `first` to the one
    containing `last`, dedented."""
    lines = open(f"{REF}/map_nav_src/r2r/agent.py").read().splitlines()
    i = next(k for k, ln in enumerate(lines) if first in ln)
    j = next(k for k in range(i, len(lines)) if last in lines[k])
    return textwrap.dedent("\n".join(lines[i:j + 1]))


@needs_ref
def test_ability_weight_draws_match_the_reference_rollout(monkeypatch):
    """MKRW (agent.py:866-869) and the 'grad' flavour (agent.py:857-860), executed from the rollout's own lines."""
    import torch.nn.functional as F
    from magic_b200 import makd
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    me = types.SimpleNamespace(args=types.SimpleNamespace(rw_temp=4.0))
    # MKRW: the reference draws torch.randn(5) from the global CPU generator, so does mkrw_weights without a generator
    code = _agent_snippet("loss_weights = torch.randn(5)", "s_softmax_weights = F.softmax(loss_weights")
    for seed in (0, 7):
        ns = dict(torch=torch, F=F, self=me)
        torch.manual_seed(seed)
        exec(code, ns)
        torch.manual_seed(seed)
        ours = makd.mkrw_weights(rw_temp=4.0)
        assert np.allclose(ours, ns["s_softmax_weights"].tolist(), rtol=1e-6) and abs(sum(ours) - 5) < 1e-5
    # 'grad': softmax(-grads / temp) * 5 in the key order [txt, img, local, global, action]
    code = _agent_snippet("key_order = ['txt', 'img', 'local', 'global', 'action']", "s_softmax_weights = F.softmax(current_iter_grads")
    grads = {"txt": 0.8, "img": 2.5, "local": 0.1, "global": 1.7, "action": 0.4}
    ns = dict(torch=torch, F=F, self=me, current_iter_grads=dict(grads))
    exec(code, ns)
    assert np.allclose(makd.grad_weights(grads, rw_temp=4.0), ns["s_softmax_weights"].tolist(), rtol=1e-6)


@needs_ref
def test_prepared_batches_travel_through_the_reference_prefetch_loader():
    """A collate batch with our index tables attached (graph_index.prepare_batch, done in the loader worker) survives
    the reference's own `move_to_cuda` / `PrefetchLoader` (data/loader.py:76-124) unchanged: tensors moved, the nested
    table dict, python lists of viewpoint ids and None entries kept."""
    from magic_b200 import synth
    from magic_b200.graph_index import prepare_batch
    ref = _load(f"{REF}/pretrain_src/data/loader.py", "ref_loader2")
    batches = [prepare_batch(synth.make_batch(t, 3, seed=5 + i)) for i, t in enumerate(("mlm", "sap"))]
    out = list(ref.PrefetchLoader(batches, torch.device("cpu")))
    assert len(out) == 2
    K = magic_b200.INDEX_KEY
    for a, b in zip(batches, out):
        assert set(a) == set(b) and isinstance(b[K], dict) and set(a[K]) == set(b[K])
        for k, v in a.items():
            if torch.is_tensor(v):
                assert torch.equal(v, b[k]), k
            elif k != K:
                assert v == b[k], k
        for k, v in a[K].items():
            assert torch.equal(v, b[K][k]) if torch.is_tensor(v) else v == b[K][k], k
