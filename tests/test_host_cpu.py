"""CPU: host-side logic of the product package, the C-ABI surface and the drop-in contracts."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import vimoclip_b200 as vmc
from oracle import indexing as oidx, student as ostudent, tfam as otfam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vimoclip_b200.h")).read()
    declared = set(re.findall(r"\b(vmc_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(vmc._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(vmc._lib.EXPORTED_SYMBOLS), declared ^ set(vmc._lib.EXPORTED_SYMBOLS)
    assert vmc._lib.lib().vmc_abi_version() == vmc._lib.ABI_VERSION == 3


def test_no_cpu_fallback():
    with pytest.raises(vmc._lib.VmcError):
        vmc.ops.prologue(torch.zeros(1, 3, 224, 224, dtype=torch.uint8), wrap=True, dst="u8")
    with pytest.raises(vmc._lib.VmcError):
        vmc.ops.gemm(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))
    tower = vmc.VisionTower(32, 128, 1, 2, 64)
    with pytest.raises(vmc._lib.VmcError):
        tower(torch.zeros(1, 3, 224, 224))
    with pytest.raises(vmc._lib.VmcError):
        vmc.AMO_CLIP(device="cpu").eval()(torch.zeros(1, 4, 512), torch.zeros(1, 4, 512))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vimo-clip_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
            assert "/root/reference" not in src, fn


def test_state_dict_layouts_match_reference_names():
    ours = vmc.FlowStudentModel("ViT-B/32", device="cpu")
    ref = ostudent.StudentOracle("ViT-B/32")
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    ours.load_state_dict(ref.state_dict(), strict=True)
    # DataParallel-style "module." checkpoints (train.py:167) load through the wrapper
    wrapped = torch.nn.DataParallel(vmc.FrameDiffStudentModel("ViT-B/32", device="cpu"))
    wrapped.load_state_dict({"module." + k: v for k, v in ref.state_dict().items()}, strict=True)
    t_ours = vmc.AMO_CLIP(device="cpu")
    t_ref = otfam.TfamOracle()
    assert {k: tuple(v.shape) for k, v in t_ours.state_dict().items()} == {k: tuple(v.shape) for k, v in t_ref.state_dict().items()}
    t_ours.load_state_dict(t_ref.state_dict(), strict=True)
    assert ours.visual_encoder.output_dim == 512 and len(ours.preprocess.transforms) == 5


def test_reference_module_paths_resolve():
    from models.student_model import FlowStudentModel
    from models.student_model_frame_diff import FrameDiffStudentModel
    from TFAM.models import AMO_CLIP
    import losses

    assert FlowStudentModel is vmc.FlowStudentModel and FrameDiffStudentModel is vmc.FrameDiffStudentModel
    assert AMO_CLIP is vmc.AMO_CLIP and losses.distillation_loss is vmc.distillation_loss


def test_indexing_bit_exact_with_oracle(golden):
    g = golden("indexing.npz")
    for key in g.files:
        if key.startswith("sparse_"):
            _, T, n = key.split("_")
            emb = torch.arange(int(T), dtype=torch.float32)[:, None]
            assert np.array_equal(vmc.indexing.sparse_sampling(emb, int(n))[:, 0].long().numpy(), g[key])
    batch = [{"video_id": str(i), "embeddings": torch.from_numpy(g[f"collate_in_rgb{i}"]),
              "flow_embeddings": torch.from_numpy(g[f"collate_in_flow{i}"]), "labels": torch.zeros(4)} for i in range(3)]
    col = vmc.indexing.collate_fn_pad(batch)
    assert np.array_equal(col["embeddings"].numpy(), g["collate_rgb"])
    assert np.array_equal(col["mask_rgb"].numpy(), g["collate_mask_rgb"])
    assert np.array_equal(col["mask_flow"].numpy(), g["collate_mask_flow"])
    for total, mx in [(10, None), (10, 16), (100, 16), (451, 32), (33, 32)]:
        assert np.array_equal(vmc.indexing.sample_frame_indices(total, mx), oidx.sample_frame_indices(total, mx))
    assert vmc.indexing.segment_frame_indices(7, 3) == oidx.segment_indices(7, 3)


def test_shard_index_math_round_trips():
    for n, w in [(256, 8), (10, 4), (7, 2), (3, 8), (1, 1)]:
        per = vmc.indexing.padded_per_rank(n, w)
        cat = np.full(w * per, -1)
        for r in range(w):
            ids = vmc.indexing.shard_ids(n, r, w)
            assert np.array_equal(ids, oidx.shard_clips(n, r, w)[0])
            cat[r * per: r * per + len(ids)] = ids
        assert np.array_equal(cat[vmc.indexing.unshard_order(n, w)], np.arange(n))


def test_losses_cpu_semantics(golden):
    g = golden("losses.npz")
    s, t = torch.from_numpy(g["s"]), torch.from_numpy(g["t"])
    assert np.isclose(float(vmc.distillation_loss(s, t, "cosine")), float(g["cos"]), atol=1e-6)
    assert np.isclose(float(vmc.distillation_loss(s, t, "mse")), float(g["mse"]), atol=1e-6)
    lg, tg = torch.from_numpy(g["logits"]), torch.from_numpy(g["targets"])
    assert np.isclose(float(vmc.classification_loss(lg, tg, positive_weight=3)), float(g["bce_pw"]), atol=1e-6)
    with pytest.raises(ValueError):
        vmc.distillation_loss(s, t, "l1")


def test_gather_clips_world2_gloo():
    """N>1 host path on CPU: two gloo ranks shard 7 clips r::2 and gather them back in order."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import vimoclip_b200 as vmc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
n = 7
ids = vmc.sharding.local_clip_ids(n)
local = torch.tensor(ids, dtype=torch.float32)[:, None].repeat(1, 3) * 10 + torch.arange(3)
out = vmc.sharding.gather_clips(local, n)
exp = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3) * 10 + torch.arange(3)
assert torch.equal(out, exp), out
dist.destroy_process_group()
print("ok")
''' % ROOT
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(os.environ, RANK=str(r), PORT=str(port)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0 and b"ok" in out, out.decode()


def test_fast_normalise_constants():
    """The gather prologue emits bf16(fma(u, K_c, B_c)) (csrc/elementwise.cu: c_fastK / c_fastB).  For all 768
    (u, c) pairs that equals bf16(RN((u/255 - mean_c)/std_c)), the exact fp32 chain of the oracle."""
    from oracle import prologue as oprol

    src = open(os.path.join(ROOT, "vimo-clip_b200", "csrc", "elementwise.cu")).read()
    ks = [int(v, 16) for v in re.search(r"c_fastK\[3\] = \{([^}]*)\}", src).group(1).replace("u", "").split(",")]
    bs = [int(v, 16) for v in re.search(r"c_fastB\[3\] = \{([^}]*)\}", src).group(1).replace("u", "").split(",")]
    u8 = np.arange(256, dtype=np.uint8).reshape(1, 1, 16, 16).repeat(3, axis=1)
    exact = torch.from_numpy(oprol.normalise_u8(u8)).to(torch.bfloat16).view(torch.int16).numpy()
    for c in range(3):
        k = float(np.uint32(ks[c]).view(np.float32))
        b = float(np.uint32(bs[c]).view(np.float32))
        # u*K is exact in float64 (8 + 24 bits) and so is the sum: one rounding to fp32 == fma semantics
        fused = (np.arange(256, dtype=np.float64) * k + b).astype(np.float32)
        got = torch.from_numpy(fused).to(torch.bfloat16).view(torch.int16).numpy()
        assert np.array_equal(got, exact[0, c].reshape(-1)), c


def test_embedding_store_layout_and_roundtrip(tmp_path):
    """The on-disk contract between the stages (extract_embeddings.py:50-55,106-119; TFAM/data/dataset.py:25-73) on the
    sidecar format: same hierarchy, attributes and h5py-style access."""
    from vimoclip_b200.store import EmbeddingStore, write_video

    path = tmp_path / "ak_val_clip_embeddings.vmc"
    rng = np.random.default_rng(0)
    embs = {f"vid{i}": rng.standard_normal((5 + i, 512)).astype(np.float32) for i in range(3)}
    with EmbeddingStore(path, "w") as hf:
        hf.attrs["num_classes"] = 140
        hf.attrs["dataset_name"] = "AnimalKingdom"
        hf.attrs["clip_model"] = "ViT-B/16"
        for vid, e in embs.items():
            lab = np.zeros(140, dtype=np.float32)
            lab[[3, 77]] = 1.0
            g = write_video(hf, vid, torch.from_numpy(e), lab, total_frames=e.shape[0], original_frames=10 * e.shape[0])
            assert g["embeddings"].shape == e.shape
        hf.create_dataset("video_ids", data=np.array(list(embs), dtype=object))
        with pytest.raises(ValueError):
            hf.create_group("vid0")  # h5py raises on duplicates too
    with EmbeddingStore(path, "r") as f:
        assert sorted(f.keys()) == ["vid0", "vid1", "vid2", "video_ids"]
        assert f.attrs["num_classes"] == 140 and f.attrs["clip_model"] == "ViT-B/16"
        assert list(f["video_ids"][:]) == list(embs)
        for vid, e in embs.items():
            assert f[vid]["embeddings"].shape[0] == e.shape[0]  # the max_frames filter of TFAM/data/dataset.py:30
            assert np.array_equal(f[vid]["embeddings"][:], e) and f[vid]["embeddings"].dtype == np.float32
            assert f[vid]["labels"][:].sum() == 2.0
            assert f[vid].attrs["total_frames"] == e.shape[0] and f[vid].attrs["original_frames"] == 10 * e.shape[0]
        with pytest.raises(OSError):
            f.create_group("nope")
        with pytest.raises(KeyError):
            f["missing"]
    # resume (inference_frame_diff.py: skip videos already present) and MammalNet-style nesting + extendable datasets
    with EmbeddingStore(path, "a") as f:
        assert "vid1" in f and "vid9" not in f
        g = f.create_group("trimmed_videos/clipA")
        ds = g.create_dataset("embeddings", shape=(0, 512), maxshape=(None, 512), dtype=np.float32, compression="gzip", chunks=(2048, 512))
        off = 0
        for chunk in (rng.standard_normal((4, 512)).astype(np.float32), rng.standard_normal((3, 512)).astype(np.float32)):
            ds.resize(off + chunk.shape[0], axis=0)  # extract_embeddings_mammalNet.py:137-141
            ds[off:off + chunk.shape[0]] = chunk
            off += chunk.shape[0]
        last = chunk
    with EmbeddingStore(path, "r") as f:
        assert "trimmed_videos" in f and list(f["trimmed_videos"].keys()) == ["clipA"]
        e = f["trimmed_videos"]["clipA"]["embeddings"]
        assert e.shape == (7, 512) and np.array_equal(e[4:], last)
        assert f["trimmed_videos/clipA"]["embeddings"].shape == (7, 512)
        f.to_hdf5(str(tmp_path / "x.h5"))  # the reference's HDF5 layout (built-in writer: h5py is absent here)
    from vimoclip_b200.hdf5_min import read_hdf5

    tree = read_hdf5(str(tmp_path / "x.h5"))
    assert tree["attrs"] == {"num_classes": 140, "dataset_name": "AnimalKingdom", "clip_model": "ViT-B/16"}
    assert sorted(tree["children"]) == ["trimmed_videos", "vid0", "vid1", "vid2", "video_ids"]
    assert list(tree["children"]["video_ids"][1]) == list(embs)  # variable-length UTF-8 strings, as h5py.string_dtype()
    for vid, e in embs.items():
        g = tree["children"][vid]
        assert np.array_equal(g["children"]["embeddings"][1], e) and g["children"]["embeddings"][1].dtype == np.float32
        assert g["attrs"] == {"total_frames": e.shape[0], "original_frames": 10 * e.shape[0]}
    back = EmbeddingStore.from_hdf5(str(tmp_path / "x.h5"), tmp_path / "back.vmc")
    assert sorted(back.keys()) == sorted(tree["children"]) and np.array_equal(back["vid1"]["embeddings"][:], embs["vid1"])
    assert back["trimmed_videos/clipA"]["embeddings"].shape == (7, 512) and back.attrs["num_classes"] == 140


def test_tfam_weight_stream_packing_matches_mma_fragment_order():
    """The fused TFAM kernel's weight stream (vimoclip_b200.tfam.pack_weight_stream): replaying, per (CTA, warp), the 512-byte
    blocks in stream order with the mma.m16n8k16 B-fragment lane mapping reproduces x @ W_slice^T for every projection."""
    import numpy as np

    from vimoclip_b200 import tfam

    torch.manual_seed(0)
    m = tfam.AMO_CLIP(device="cpu", num_layers=2)
    ws = tfam.pack_weight_stream(m.layers)
    assert tuple(ws.shape) == (8, 8, 2, 256, 32, 8) and ws.dtype == torch.float16
    lay = {"sin": (0, 3, 16), "sout": (48, 1, 16), "cq": (64, 1, 16), "ckv": (80, 2, 16), "cout": (112, 1, 16), "f1": (128, 4, 16),
           "f2": (192, 8, 8)}
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    koff = np.stack([2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9, 16 + 2 * t, 17 + 2 * t, 24 + 2 * t, 25 + 2 * t], 1)  # [32, 8]

    def replay(a, blocks, nt, kp):
        out = np.zeros((a.shape[0], nt * 8))
        for b in range(nt * kp):
            p, i = divmod(b, nt)
            blk = blocks[b].astype(np.float64)  # [32, 8]
            for ln in range(32):
                out[:, 8 * i + g[ln]] += a[:, 32 * p + koff[ln]] @ blk[ln]
        return out

    rs = np.random.RandomState(0)
    a512, a256 = rs.randn(4, 512), rs.randn(4, 256)
    w16 = lambda w: w.detach().to(torch.float16).double().numpy()  # noqa: E731
    ly, layer = m.layers[1], 1
    for c, w in [(0, 0), (5, 2), (7, 7)]:
        st = ws[c, w, layer].numpy()
        take = lambda k: st[lay[k][0]:lay[k][0] + lay[k][1] * lay[k][2]]  # noqa: E731
        r0 = 64 * c + 8 * w
        wi, wc = w16(ly.self_attn.in_proj_weight), w16(ly.cross_attn.in_proj_weight)
        want = {
            "sin": np.concatenate([a512 @ wi[512 * i + r0:512 * i + r0 + 8].T for i in range(3)], 1),
            "sout": a512 @ w16(ly.self_attn.out_proj.weight)[r0:r0 + 8].T,
            "cq": a512 @ wc[r0:r0 + 8].T,
            "ckv": np.concatenate([a512 @ wc[512 * (i + 1) + r0:512 * (i + 1) + r0 + 8].T for i in range(2)], 1),
            "cout": a512 @ w16(ly.cross_attn.out_proj.weight)[r0:r0 + 8].T,
            "f1": np.concatenate([a512 @ w16(ly.ffn[0].weight)[256 * c + 32 * w + 8 * i:256 * c + 32 * w + 8 * i + 8].T for i in range(4)], 1),
            "f2": np.concatenate([a256 @ w16(ly.ffn[3].weight)[64 * w + 8 * i:64 * w + 8 * i + 8, 256 * c:256 * c + 256].T for i in range(8)], 1),
        }
        for k, (_, nt, kp) in lay.items():
            got = replay(a256 if k == "f2" else a512, take(k), nt, kp)
            assert np.abs(got - want[k]).max() < 1e-9, (c, w, k)


def test_embedding_store_flush_per_video_is_journalled_and_crash_safe(tmp_path):
    """inference_frame_diff.py:299,395,404 flushes after EVERY video and opens with h5py keyword arguments: each flush appends
    only the change to journal.jsonl (not the whole index), a reader that arrives after a crash (no close()) sees every
    flushed video, close() compacts, and mode 'w' removes only the files the old index lists."""
    import json

    from vimoclip_b200.store import EmbeddingStore, write_video

    path = tmp_path / "frame_diff_embeddings.vmc"
    keep = tmp_path / "frame_diff_embeddings.vmc" / "dont_touch.npy"
    rng = np.random.default_rng(1)
    hf = EmbeddingStore(path, "a", libver="latest")  # h5py.File(path, 'a', libver='latest')
    np.save(keep, np.arange(3))  # a foreign file whose name looks like ours
    sizes = []
    for i in range(40):
        write_video(hf, f"v{i:03d}", rng.standard_normal((3, 8)).astype(np.float32), np.zeros(5, np.float32), 3, 30)
        hf.flush()
        sizes.append(os.path.getsize(path / "journal.jsonl"))
    growth = np.diff(sizes)
    assert growth.max() < 2 * growth.min() + 64  # every flush appends O(1) bytes: no O(N) index rewrite per video
    # "crash": no close().  A new reader replays snapshot + journal.
    with EmbeddingStore(path, "r") as f:
        assert len(f.keys()) == 40 and f["v039"].attrs["original_frames"] == 30
        assert f["v017"]["embeddings"].shape == (3, 8)
    hf.close()
    assert not os.path.exists(path / "journal.jsonl")
    with open(path / "index.json") as fh:
        assert len(json.load(fh)["datasets"]) == 80
    with EmbeddingStore(path, "w") as f:  # truncate: our datasets go, the foreign file stays
        assert len(f.keys()) == 0
    assert os.path.exists(keep) and len([n for n in os.listdir(path) if n.endswith(".npy")]) == 1


def test_hdf5_reader_on_a_libhdf5_written_file_and_writer_round_trip(tmp_path):
    """vimoclip_b200.hdf5_min: (i) the reader parses a GENUINE libhdf5-written file -- scipy ships a MATLAB v7.3 file (HDF5 behind
    a 512-byte user block) whose content is known: ``testdouble = 0:pi/4:2*pi`` with the attribute MATLAB_class = 'double';
    (ii) files written here (30 k-video scale structures in miniature: multi-level group B-trees, chunked + gzip embeddings
    with a two-level chunk B-tree, variable-length strings) read back identically; (iii) libver='latest' files are refused."""
    import scipy.io.matlab

    from vimoclip_b200.hdf5_min import Hdf5FormatError, read_hdf5, write_hdf5

    real = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if os.path.exists(real):
        t = read_hdf5(real)
        kind, arr, attrs = t["children"]["testdouble"]
        assert arr.shape == (9, 1) and arr.dtype == np.float64 and attrs == {"MATLAB_class": "double"}
        assert np.allclose(arr.ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    rng = np.random.default_rng(0)
    root = {"attrs": {"num_classes": 140, "clip_model": "ViT-B/16", "scale": 0.5}, "children": {}}
    for i in range(300):  # > 8 * 32 links: a two-level group B-tree
        emb = rng.standard_normal((3 + i % 5, 64)).astype(np.float32)
        root["children"][f"vid{i:04d}"] = {"attrs": {"total_frames": emb.shape[0], "original_frames": 10 * i}, "children": {
            "embeddings": ("dataset", emb, {}, {"chunks": (1, 64), "compression": "gzip"}),
            "labels": ("dataset", (rng.random(140) < 0.05).astype(np.float32), {})}}
    long_emb = rng.standard_normal((150, 64)).astype(np.float32)  # 150 chunks > 64: a two-level chunk B-tree
    root["children"]["long"] = {"attrs": {}, "children": {"embeddings": ("dataset", long_emb, {}, {"chunks": (1, 64), "compression": "gzip"})}}
    root["children"]["video_ids"] = ("dataset", [f"vid{i:04d}" for i in range(300)], {})
    path = str(tmp_path / "ak.h5")
    write_hdf5(path, root)
    back = read_hdf5(path)
    assert back["attrs"] == root["attrs"] and sorted(back["children"]) == sorted(root["children"])
    for k, v in root["children"].items():
        if isinstance(v, dict) and k.startswith("vid"):
            assert back["children"][k]["attrs"] == v["attrs"]
            for leaf in ("embeddings", "labels"):
                assert np.array_equal(back["children"][k]["children"][leaf][1], v["children"][leaf][1])
    assert np.array_equal(back["children"]["long"]["children"]["embeddings"][1], long_emb)
    assert list(back["children"]["video_ids"][1]) == [f"vid{i:04d}" for i in range(300)]
    data = bytearray(open(path, "rb").read())
    assert data[:8] == b"\x89HDF\r\n\x1a\n" and len(data) % 8 == 0
    data[8] = 2  # a superblock version this reader does not cover
    bad = str(tmp_path / "latest.h5")
    open(bad, "wb").write(bytes(data))
    with pytest.raises(Hdf5FormatError):
        read_hdf5(bad)
