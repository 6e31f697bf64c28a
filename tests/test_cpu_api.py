"""CPU-side checks: the package's CPU route against the golden vectors and the live reference (bitwise), the
state_dict/API surface, the C-ABI library's exported symbols, and the reference's own tests run against this package."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import _golden
import _live_reference as live

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-rechub_b200")


@pytest.mark.parametrize("name", _golden.NAMES)
def test_package_cpu_route_matches_reference_golden(name):
    import torch_rechub.basic.features as F
    import torch_rechub.models.ranking as M
    rec = _golden.load(name)
    model = _golden.build_model(name, rec, F, M)
    assert list(model.state_dict().keys()) == list(rec["sd"].keys())  # checkpoint layout = the reference's
    x, y = _golden.torch_inputs(rec)
    model.eval()  # first: the golden state_dict holds the running statistics the reference's eval pass used
    with torch.no_grad():
        assert np.abs(model(x).numpy() - rec["eval_prob"]).max() <= 1e-7
    model.train()
    p = model(x)
    assert np.abs(p.detach().numpy() - rec["train_prob"]).max() <= 1e-7
    torch.nn.BCELoss()(p, y).backward()
    for k, prm in model.named_parameters():
        ref = rec["grad"][k]
        scale = max(np.abs(ref).max(), 1e-3)
        if k.endswith(".bias") and k[:-4] + "weight" in rec["grad"]:
            scale = max(scale, np.abs(rec["grad"][k[:-4] + "weight"]).max())
        assert np.abs(prm.grad.numpy() - ref).max() <= 1e-5 * scale, k


@pytest.mark.skipif(not live.live_reference_available(), reason="live reference only exists in the build container")
def test_same_seed_same_weights_same_outputs_as_live_reference():
    """Initialisers consume the RNG like the reference (initializers.py:16-21): same seed -> identical tables, and the
    CPU route is the same op sequence -> bitwise identical probabilities and (DeepFM/DCN/WideDeep) gradients."""
    import torch_rechub.basic.features as F
    import torch_rechub.models.ranking as M
    RF, RM = live.ref_module("basic.features"), live.ref_module("models.ranking")

    def build(mf, mm):
        torch.manual_seed(7)
        dense = [mf.DenseFeature("I%d" % i) for i in range(3)]
        sparse = [mf.SparseFeature("C%d" % i, vocab_size=50 + i, embed_dim=8) for i in range(4)] + [mf.SparseFeature("P", vocab_size=9, embed_dim=8, padding_idx=0)]
        return [
            mm.DeepFM(dense + sparse, sparse, {"dims": [16, 8], "dropout": 0.0, "activation": "relu"}),
            mm.DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [16, 8]}),
            mm.WideDeep(dense, sparse, {"dims": [8]}),
        ]

    g = torch.Generator().manual_seed(0)
    x = {"I%d" % i: torch.rand(32, generator=g) for i in range(3)}
    x.update({"C%d" % i: torch.randint(0, 50, (32,), generator=g) for i in range(4)})
    x["P"] = torch.randint(0, 9, (32,), generator=g)
    for a, b in zip(build(F, M), build(RF, RM)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
    # gradients: fresh models (the feature objects above share tables across the three models of one build)
    for idx in range(3):
        a, b = build(F, M)[idx], build(RF, RM)[idx]
        ya, yb = a(x), b(x)
        assert torch.equal(ya, yb)
        ya.sum().backward()
        yb.sum().backward()
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            assert torch.equal(p.grad, q.grad), n
    # feature objects reused across models share ONE table
    from torch_rechub.basic.layers import EmbeddingLayer
    f = F.SparseFeature("z", 10, 4)
    assert EmbeddingLayer([f]).embed_dict["z"] is EmbeddingLayer([f]).embed_dict["z"] is f.embed


def test_layer_contracts():
    from torch_rechub.basic.features import DenseFeature, SequenceFeature, SparseFeature
    from torch_rechub.basic.layers import EmbeddingLayer, InputMask
    from torch_rechub.utils.data import get_auto_embedding_dim
    assert SparseFeature("a", 10000).embed_dim == get_auto_embedding_dim(10000) == 60
    assert repr(SparseFeature("a", 10, 4)) == "<SparseFeature a with Embedding shape (10, 4)>"
    assert repr(SequenceFeature("s", 10, 4)) == "<SequenceFeature s with Embedding shape (10, 4)>"
    assert repr(DenseFeature("d")) == "<DenseFeature d>"
    feats = [SparseFeature("a", 10, 4), DenseFeature("d"), SparseFeature("b", 10, 4, shared_with="a")]
    layer = EmbeddingLayer(feats)
    assert list(layer.embed_dict.keys()) == ["a"] and layer.n_dense == 1
    assert isinstance(layer.embed_dict["a"], torch.nn.Embedding)
    x = {"a": torch.tensor([1, 2]), "b": torch.tensor([2.9, 1.0]), "d": torch.tensor([0.5, 0.25], dtype=torch.float64)}
    out = layer(x, feats, squeeze_dim=True)
    assert out.shape == (2, 9) and out.dtype == torch.float32
    w = layer.embed_dict["a"].weight
    assert torch.equal(out[0], torch.cat([w[1], w[2], torch.tensor([0.5])]))  # sparse first, dense appended; float ids truncate
    assert layer(x, feats).shape == (2, 2, 4)
    with pytest.raises(ValueError):
        layer(x, [feats[1]], squeeze_dim=False)
    with pytest.raises(ValueError):
        InputMask()(x, [feats[1]])
    with pytest.raises(IndexError):
        layer({"a": torch.tensor([10])}, [feats[0]])
    bad = SequenceFeature("s", 5, 4, pooling="max")
    with pytest.raises(ValueError):
        EmbeddingLayer([bad])({"s": torch.zeros(1, 2).long()}, [bad])


def test_out_of_scope_models_are_importable_but_refuse_to_build():
    from torch_rechub.models.ranking import EDCN, AutoInt, DeepFFM, FatDeepFFM, FiBiNet  # run_criteo.py:10 imports these
    for cls in (EDCN, AutoInt, DeepFFM, FatDeepFFM, FiBiNet):
        with pytest.raises(NotImplementedError):
            cls()


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rechub_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rh_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib_path = os.path.join(PKG, "lib", "librechub_b200.so")
    assert os.path.exists(lib_path), "build the engine first (python -c 'import __graft_entry__ as g; g.build()')"
    from torch_rechub.b200 import _lib
    handle = _lib.lib()  # dlopen + argtypes for every prototype; no compute
    for name in declared:
        assert hasattr(handle, name), "librechub_b200.so does not export %s" % name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert handle.rh_abi_version() == 1


def test_cuda_route_has_no_silent_fallback(monkeypatch):
    """A missing engine library must raise, not fall back to eager ops."""
    from torch_rechub.b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librechub_b200.so")
    with pytest.raises(_lib.EngineMissing):
        _lib.lib()


@pytest.mark.skipif(not live.live_reference_available(), reason="the reference's test files only exist in the build container")
def test_reference_own_tests_pass_against_this_package(tmp_path):
    """The reference's own test files, byte-for-byte, run against THIS package.  They are copied to a scratch directory
    first because they put their own parent directory (the reference checkout) at the front of sys.path."""
    import shutil
    tdir = tmp_path / "suite" / "tests"
    tdir.mkdir(parents=True)
    for name in ("test_regularization.py", "test_e2e_ranking.py", "test_parquet_dataset.py", "test_pa_array_to_tensor.py"):
        shutil.copy(os.path.join(live.REFERENCE_ROOT, "tests", name), tdir / name)
    env = dict(os.environ, PYTHONPATH=PKG)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", str(tdir), "-k", "regularization or WideDeep or (DCN and not EDCN) or parquet or pa_array or tensor or dataset or Parquet"]
    res = subprocess.run(cmd, env=env, cwd=str(tdir), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "passed" in res.stdout


def test_batched_crossnetmix_route_equals_the_reference_loop():
    """CrossNetMix's CUDA route batches the experts into three GEMMs per layer (gate-weighted sum inside the last product's K
    dimension).  Exercised here on CPU tensors against the reference-order loop: outputs and every gradient agree to fp32
    reassociation level, including the B = 1 squeeze() quirk."""
    from torch_rechub.basic.layers import CrossNetMix
    torch.manual_seed(0)
    for batch in (1, 7, 300):
        m = CrossNetMix(61, num_layers=3, low_rank=8, num_experts=3)
        with torch.no_grad():
            for b in m.bias:
                b.normal_(0, 0.1)
        x = torch.randn(batch, 61) * 0.5
        x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        y1, y2 = m(x1), m._forward_batched(x2)
        assert y1.shape == y2.shape
        assert (y1 - y2).abs().max().item() <= 1e-6 * max(1.0, y1.abs().max().item())
        g = torch.randn_like(y1)
        (y1 * g).sum().backward()
        ref = {n: p.grad.clone() for n, p in m.named_parameters()}
        m.zero_grad()
        (y2 * g).sum().backward()
        for n, p in m.named_parameters():
            assert (p.grad - ref[n]).abs().max().item() <= 2e-5 * ref[n].abs().max().item() + 1e-7, n
        assert (x1.grad - x2.grad).abs().max().item() <= 2e-5 * x1.grad.abs().max().item() + 1e-7


def test_packed_loader_int32_ids_feed_the_same_model_outputs():
    from torch_rechub.b200.data import PackedLoader
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.models.ranking import DeepFM
    g = np.random.RandomState(0)
    n = 50
    x = {"I0": g.rand(n), "C0": g.randint(0, 30, n), "C1": g.randint(0, 30, n)}
    y = g.randint(0, 2, n)
    torch.manual_seed(0)
    feats = [SparseFeature("C0", 30, 8), SparseFeature("C1", 30, 8)]
    model = DeepFM([DenseFeature("I0")] + feats, feats, {"dims": [8]}).eval()
    outs = []
    for dt in (torch.int64, torch.int32):
        (xb, yb), = list(PackedLoader(x, y, batch_size=n, id_dtype=dt))
        assert xb["C0"].dtype == dt and xb.h2d_bytes() == n * (2 * (8 if dt == torch.int64 else 4) + 4)
        outs.append(model(xb))
    assert torch.equal(outs[0], outs[1])
    with pytest.raises(ValueError):
        PackedLoader({"C0": np.array([2**31])}, np.array([0]), batch_size=1, id_names=["C0"], num_names=[], id_dtype=torch.int32)
    with pytest.raises(ValueError):
        PackedLoader(x, y, batch_size=n, id_dtype=torch.int16)


_EXAMPLE_RUNNER = """import random, runpy, sys
import numpy as np
random.seed(0); np.random.seed(0)  # utils/data.py:206-241 shuffles and draws negatives from the unseeded `random` module
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(script, run_name="__main__")
"""


@pytest.mark.skipif(not live.live_reference_available() or not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "torch_rechub")),
                    reason="the reference's examples and its installed copy (oracle/_ref) only exist in the build container")
@pytest.mark.parametrize("script,args", [("run_criteo.py", ["--model_name", "deepfm"]), ("run_criteo.py", ["--model_name", "dcn_v2"]), ("run_amazon_electronics.py", [])])
def test_reference_quickstart_examples_run_unchanged_and_reproduce_the_reference(tmp_path, script, args):
    """BASELINE configs[0]: the reference's own example scripts (examples/ranking/run_criteo.py on the shipped Criteo sample — DeepFM,
    DCN-v2 — and run_amazon_electronics.py — DIN), byte for byte, on CPU: once against THIS package, once against the unmodified
    reference installed under oracle/_ref.  Same seeds -> the same validation / test AUC lines, digit for digit."""
    import shutil
    ex = os.path.join(live.REFERENCE_ROOT, "examples", "ranking")
    shutil.copy(os.path.join(ex, script), tmp_path / script)
    sub = "criteo" if script == "run_criteo.py" else "amazon-electronics"
    shutil.copytree(os.path.join(ex, "data", sub), tmp_path / "data" / sub)
    (tmp_path / "runner.py").write_text(_EXAMPLE_RUNNER)
    lines = {}
    for name, path in (("package", PKG), ("reference", os.path.join(ROOT, "oracle", "_ref"))):
        env = dict(os.environ, PYTHONPATH=path)
        res = subprocess.run([sys.executable, "runner.py", script, "--epoch", "2", "--device", "cpu"] + args, env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, (name, res.stdout[-2000:] + res.stderr[-2000:])
        lines[name] = [ln.strip() for ln in res.stdout.splitlines() if "auc" in ln]
        assert any(ln.startswith("test auc:") for ln in lines[name]), (name, res.stdout[-1000:])
    assert lines["package"] == lines["reference"], lines


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The C ABI's structs (rh_field, rh_dense, rh_sync) as gcc lays them out from include/rechub_b200.h against the ctypes mirrors in
    torch_rechub.b200._lib: same field names in the same order, same offsets, same sizes (no compute, no GPU)."""
    import ctypes
    from torch_rechub.b200 import _lib
    header = open(os.path.join(ROOT, "include", "rechub_b200.h")).read()
    pairs = {"rh_field": _lib.RhField, "rh_dense": _lib.RhDense, "rh_sync": _lib.RhSync}
    prog = ["#include <stdio.h>", "#include <stddef.h>", '#include "rechub_b200.h"', "int main(void) {"]
    names = {}
    for cname in pairs:
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, flags=re.S)
        assert m, cname
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
        fields = [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]
        names[cname] = fields
        prog.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for f in fields:
            prog.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, f))
        prog.append('  printf("\\n");')
    prog += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        cname, size, offs = parts[0], int(parts[1]), [int(v) for v in parts[2:]]
        st = pairs[cname]
        assert [f[0] for f in st._fields_] == names[cname], (cname, [f[0] for f in st._fields_], names[cname])
        assert ctypes.sizeof(st) == size, (cname, ctypes.sizeof(st), size)
        assert [getattr(st, f).offset for f in names[cname]] == offs, cname


def test_packed_loader_shuffled_batches_are_contiguous_views_of_chunk_buffers():
    """the shuffled path no longer fancy-indexes into fresh pageable memory.  Every shuffled batch is a contiguous view
    of one of two chunk buffers (pinned on a CUDA box) and carries exactly the rows `randperm` assigned to it — across chunk
    boundaries, a ragged last batch, drop_last and a second epoch that reuses the buffers."""
    from torch_rechub.b200.data import PackedLoader
    n, bs = 1000, 64
    x = {"a": np.arange(n), "b": np.arange(n) * 2, "f": np.arange(n).astype(np.float32) / 7, "s": np.arange(n * 3).reshape(n, 3)}
    y = np.arange(n) % 2
    for drop in (False, True):
        loader = PackedLoader(x, y, batch_size=bs, shuffle=True, drop_last=drop)
        loader._CHUNK_BATCHES = 3  # several chunk refills per epoch
        for epoch in range(2):
            torch.manual_seed(5 + epoch)
            order = torch.randperm(n)
            torch.manual_seed(5 + epoch)
            rows, ptrs = 0, set()
            for bi, (xb, yb) in enumerate(loader):
                idx = order[bi * bs:min((bi + 1) * bs, n)]
                assert torch.equal(xb["a"], torch.from_numpy(x["a"])[idx]) and torch.equal(xb["b"], torch.from_numpy(x["b"])[idx])
                assert torch.allclose(xb["f"], torch.from_numpy(x["f"])[idx]) and torch.equal(xb["s"], torch.from_numpy(x["s"])[idx])
                assert torch.equal(yb, torch.from_numpy(y).float()[idx])
                assert xb.ids.is_contiguous() and xb.nums.is_contiguous() and xb.seqs.is_contiguous()
                ptrs.add(xb.ids.untyped_storage().data_ptr())
                rows += len(idx)
            assert rows == (n // bs * bs if drop else n)
            assert len(ptrs) <= 2  # two chunk buffers, no per-batch allocation


def test_bench_region_repetition_count_is_a_function_of_the_first_region_only():
    """bench.py repeats its K-step timed region until ~60 ms are covered (3..15 regions) and reports the median; the count must follow
    from the first region's (all-reduced) time alone, so that every rank of a multi-GPU run executes the same number of regions."""
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    for first_ms, expect in ((3.3, 15), (33.0, 3), (25.0, 3), (10.0, 6), (61.0, 1), (0.0, 15)):
        calls = []

        def region(r, first_ms=first_ms):
            calls.append(r)
            return (first_ms if r == 0 else 1e9), r  # later regions (however long) never change the count

        out, last = bench.repeat_regions(region)
        assert calls == list(range(expect)) and len(out) == expect and last == expect - 1
