"""Embedding store: the on-disk contract between the three stages of the reference, without libhdf5.

The reference hands embeddings from stage to stage through HDF5 files: ``extract_embeddings.py:50-55,106-119`` writes one
group per ``video_id`` with ``embeddings [T, D]`` float32 (gzip, chunks ``(1, D)``), ``labels [C]`` float32 multi-hot, the
attributes ``total_frames`` / ``original_frames``, root attributes and a root ``video_ids`` string dataset;
``extract_embeddings_mammalNet.py:114-141`` grows ``embeddings`` chunk by chunk (``resize``); ``inference.py:94-112`` /
``inference_frame_diff.py:235-312`` write the student's embeddings the same way (resumably); ``TFAM/data/dataset.py:25-73``
reads ``f.keys()``, ``f[key]["embeddings"].shape`` / ``[:]`` and ``f[key]["labels"][:]``; MammalNet nests the groups under
``trimmed_videos/`` (``dataset_frame_diff_mn.py:42``).

No HDF5 library exists in this image (``h5py``, ``tables`` and ``libhdf5`` are all absent), so ``EmbeddingStore`` keeps the SAME
logical layout and the subset of the ``h5py.File`` API those call sites use, on a working format made for resumable appends: a
directory with one ``.npy`` per dataset (memory-mappable, written atomically) and ``index.json`` + ``journal.jsonl`` for the
hierarchy and attributes.  ``to_hdf5`` / ``from_hdf5`` convert to and from the reference's exact HDF5 files -- through h5py where it
is installed, otherwise through the built-in writer / reader of ``hdf5_min`` (specification-following, pinned against a
libhdf5-written file found in the image).  Host-side I/O only: nothing here touches the GPU path.
"""
from __future__ import annotations

import json
import os
import tempfile

import numpy as np

_INDEX = "index.json"
_JOURNAL = "journal.jsonl"  # index changes since the last snapshot, one JSON object per flush()


class _Attrs(dict):
    def __init__(self, owner, group, *a):
        super().__init__(*a)
        self._owner, self._group = owner, group

    def __setitem__(self, k, v):
        if isinstance(v, (np.generic,)):
            v = v.item()
        elif isinstance(v, np.ndarray):
            v = v.item() if v.ndim == 0 else v.tolist()
        super().__setitem__(k, v)
        self._owner._dirty_groups.add(self._group)
        self._owner._dirty()


class Dataset:
    """``h5py.Dataset`` subset: ``shape``, ``dtype``, ``[...]`` read / slice write, ``resize(n, axis=0)``."""

    def __init__(self, store, name, meta):
        self._store, self.name, self._meta = store, name, meta
        self._data = None  # loaded / pending array

    def _load(self):
        if self._data is None:
            if self._meta["kind"] == "str":
                with open(os.path.join(self._store.path, self._meta["file"]), encoding="utf-8") as f:
                    self._data = np.array(json.load(f), dtype=object)
            else:
                self._data = np.load(os.path.join(self._store.path, self._meta["file"]), mmap_mode="r" if self._store.mode == "r" else None)
        return self._data

    @property
    def shape(self):
        return tuple(self._meta["shape"])

    @property
    def dtype(self):
        return np.dtype(object) if self._meta["kind"] == "str" else np.dtype(self._meta["dtype"])

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        out = self._load()[key]
        return np.array(out) if isinstance(out, np.memmap) else out

    def _writable(self):
        arr = self._load()
        if isinstance(arr, np.memmap) or not arr.flags.writeable or arr.base is None and not arr.flags.owndata:
            arr = np.array(arr)
            self._data = arr
        return arr

    def __setitem__(self, key, value):
        self._store._need_write()
        self._writable()[key] = value  # in place: the chunk loop of extract_embeddings_mammalNet.py:137-141 stays linear
        self._store._pending[self.name] = self

    def resize(self, size, axis=0):
        """Grow (or shrink) along ``axis`` like an extendable HDF5 dataset (extract_embeddings_mammalNet.py:137-141).  Growth
        along axis 0 doubles a hidden capacity buffer, so appending chunk after chunk copies O(N) bytes in total."""
        self._store._need_write()
        arr = self._writable()
        shape = list(arr.shape)
        shape[axis] = int(size)
        cap = getattr(self, "_cap", None)
        if axis == 0 and cap is not None and cap.shape[1:] == arr.shape[1:] and cap.shape[0] >= shape[0] and arr.base is cap:
            if shape[0] > arr.shape[0]:
                cap[arr.shape[0]:shape[0]] = 0
            new = cap[:shape[0]]
        else:
            grow = list(shape)
            if axis == 0:
                grow[0] = max(shape[0], 2 * arr.shape[0])
            self._cap = np.zeros(grow, dtype=arr.dtype)
            sl = tuple(slice(0, min(a, b)) for a, b in zip(arr.shape, shape))
            self._cap[sl] = arr[sl]
            new = self._cap[tuple(slice(0, n) for n in shape)]
        self._data = new
        self._meta["shape"] = shape
        self._store._pending[self.name] = self
        self._store._dirty_ds.add(self.name)
        self._store._dirty()


class Group:
    """``h5py.Group`` subset: ``create_group``, ``create_dataset``, ``keys``, ``[...]``, ``in``, ``attrs``."""

    def __init__(self, store, name):
        self._store, self.name = store, name

    def _full(self, child):
        return child if not self.name else f"{self.name}/{child}"

    @property
    def attrs(self):
        return self._store._attrs_of(self.name)

    def keys(self):
        pre = self.name + "/" if self.name else ""
        names = [n for n in list(self._store._index["groups"]) + list(self._store._index["datasets"]) if n and n.startswith(pre)]
        return sorted({n[len(pre):].split("/")[0] for n in names})

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def __contains__(self, child):
        full = self._full(child)
        return full in self._store._index["groups"] or full in self._store._index["datasets"]

    def __getitem__(self, child):
        full = self._full(child)
        if full in self._store._index["datasets"]:
            return self._store._dataset(full)
        if full in self._store._index["groups"]:
            return Group(self._store, full)
        raise KeyError(f"unable to open object '{full}' (it does not exist)")

    def create_group(self, child):
        self._store._need_write()
        full = self._full(child)
        if full in self._store._index["groups"] or full in self._store._index["datasets"]:
            raise ValueError(f"unable to create group (name '{full}' already exists)")  # h5py raises ValueError too
        parts = full.split("/")
        for i in range(1, len(parts) + 1):  # intermediate groups, as h5py creates them
            name = "/".join(parts[:i])
            if name not in self._store._index["groups"]:
                self._store._index["groups"][name] = {}
                self._store._dirty_groups.add(name)
        self._store._dirty()
        return Group(self._store, full)

    def require_group(self, child):
        return self[child] if child in self else self.create_group(child)

    def create_dataset(self, child, data=None, shape=None, dtype=None, maxshape=None, compression=None, chunks=None, **_kw):
        """``compression`` / ``chunks`` / ``maxshape`` are accepted and recorded (they matter only for the HDF5 export)."""
        self._store._need_write()
        full = self._full(child)
        if full in self._store._index["datasets"] or full in self._store._index["groups"]:
            raise ValueError(f"unable to create dataset (name '{full}' already exists)")
        if "/" in full:
            parent = full.rsplit("/", 1)[0]
            if parent not in self._store._index["groups"]:
                Group(self._store, "").create_group(parent)
        if data is None:
            data = np.zeros(shape, dtype=dtype if dtype is not None else np.float32)
        is_str = isinstance(data, (list, tuple)) and all(isinstance(s, (str, bytes)) for s in data) or \
            (isinstance(data, np.ndarray) and data.dtype.kind in "OUS")
        if is_str:
            arr = np.array([s.decode("utf-8") if isinstance(s, bytes) else str(s) for s in np.asarray(data, dtype=object).ravel()], dtype=object)
            meta = {"kind": "str", "shape": [len(arr)], "dtype": "str"}
        else:
            arr = np.asarray(data) if dtype is None else np.asarray(data, dtype=dtype)
            meta = {"kind": "num", "shape": list(arr.shape), "dtype": arr.dtype.str}
        meta.update({"file": f"d{self._store._index['next_file']:08d}" + (".json" if is_str else ".npy"),
                     "compression": compression, "chunks": list(chunks) if chunks else None,
                     "maxshape": [None if m is None else int(m) for m in maxshape] if maxshape else None})
        self._store._index["next_file"] += 1
        self._store._index["datasets"][full] = meta
        ds = Dataset(self._store, full, meta)
        ds._data = arr
        self._store._cache[full] = ds
        self._store._pending[full] = ds
        self._store._dirty_ds.add(full)
        self._store._dirty()
        return ds


class EmbeddingStore(Group):
    """``EmbeddingStore(path, mode)`` ~ ``h5py.File(path, mode)`` for modes ``"w"``, ``"a"`` (resume) and ``"r"``.

    ``flush()`` writes pending datasets (atomic rename) and appends the index changes to ``journal.jsonl``; ``close()`` / leaving
    the ``with`` block compacts them into ``index.json``.  A crash loses at most the groups written since the last flush
    (``inference_frame_diff.py`` flushes after every video)."""

    def __init__(self, path, mode="r", **_h5py_kwargs):
        """``_h5py_kwargs`` (``libver=``, ``swmr=``, ``driver=`` ...) are accepted and ignored, so ``h5py.File(path, 'a',
        libver='latest')`` call sites (inference_frame_diff.py) carry over unchanged."""
        if mode not in ("r", "w", "a", "r+", "w-", "x"):
            raise ValueError("mode must be 'r', 'r+', 'w', 'w-' / 'x' or 'a'")
        self.path, self.mode = str(path), mode
        self._cache, self._pending, self._is_dirty, self._closed = {}, {}, False, False
        self._dirty_groups, self._dirty_ds, self._journal_lines = set(), set(), 0
        idx = os.path.join(self.path, _INDEX)
        if mode in ("w-", "x") and os.path.exists(idx):
            raise FileExistsError(f"unable to create store '{self.path}' (it exists)")
        if mode in ("r", "r+") or (mode == "a" and os.path.exists(idx)):
            if not os.path.exists(idx):
                raise FileNotFoundError(f"unable to open store '{self.path}' (no {_INDEX})")
            self._index = self._read_index()
            if mode == "r+":
                self.mode = "a"
        else:
            os.makedirs(self.path, exist_ok=True)
            if os.path.exists(idx):  # mode 'w' truncates, like h5py -- but removes only what the old index lists
                try:
                    old = self._read_index()
                    for meta in old["datasets"].values():
                        fn = os.path.join(self.path, meta["file"])
                        if os.path.exists(fn):
                            os.remove(fn)
                except (OSError, ValueError, KeyError):
                    pass
                for fn in (_INDEX, _JOURNAL):
                    if os.path.exists(os.path.join(self.path, fn)):
                        os.remove(os.path.join(self.path, fn))
            self._index = {"format": "vimoclip_b200.EmbeddingStore/1", "attrs": {}, "groups": {"": {}}, "datasets": {}, "next_file": 0}
            self._is_dirty = True
            self._snapshot()
            self.mode = "a" if mode != "w" else "w"
        super().__init__(self, "")
        self._attr_objs = {}

    def _read_index(self):
        """Snapshot + the journal of later flushes (a crash between flushes loses nothing that was flushed)."""
        with open(os.path.join(self.path, _INDEX), encoding="utf-8") as f:
            index = json.load(f)
        jp = os.path.join(self.path, _JOURNAL)
        if os.path.exists(jp):
            with open(jp, encoding="utf-8") as f:
                for line in f:
                    line = line.strip()
                    if not line:
                        continue
                    try:
                        d = json.loads(line)
                    except ValueError:
                        break  # torn last line of an interrupted flush
                    index["groups"].update(d.get("groups", {}))
                    index["datasets"].update(d.get("datasets", {}))
                    if "attrs" in d:
                        index["attrs"] = d["attrs"]
                    index["next_file"] = max(index["next_file"], d.get("next_file", 0))
                    self._journal_lines += 1
        return index

    def _snapshot(self):
        fd, tmp = tempfile.mkstemp(dir=self.path, suffix=".tmp")
        with os.fdopen(fd, "w", encoding="utf-8") as f:
            json.dump(self._index, f)
        os.replace(tmp, os.path.join(self.path, _INDEX))
        jp = os.path.join(self.path, _JOURNAL)
        if os.path.exists(jp):
            os.remove(jp)
        self._journal_lines = 0

    # -- internals --
    def _need_write(self):
        if self.mode == "r":
            raise OSError("store opened read-only")
        if self._closed:
            raise ValueError("store is closed")

    def _dirty(self):
        self._is_dirty = True

    def _attrs_of(self, group):
        if group not in self._attr_objs:
            raw = self._index["attrs"] if group == "" else self._index["groups"][group]
            self._attr_objs[group] = _Attrs(self, group, raw)
        return self._attr_objs[group]

    def _dataset(self, full):
        if full not in self._cache:
            self._cache[full] = Dataset(self, full, self._index["datasets"][full])
        return self._cache[full]

    # -- file-level API --
    def flush(self):
        if self.mode == "r" or not (self._is_dirty or self._pending):
            return
        for name, ds in list(self._pending.items()):
            target = os.path.join(self.path, ds._meta["file"])
            fd, tmp = tempfile.mkstemp(dir=self.path, suffix=".tmp")
            with os.fdopen(fd, "wb") as f:
                if ds._meta["kind"] == "str":
                    f.write(json.dumps([str(s) for s in ds._data]).encode("utf-8"))
                else:
                    np.save(f, np.ascontiguousarray(ds._data), allow_pickle=False)
            os.replace(tmp, target)
            ds._meta["shape"] = list(ds._data.shape)
        self._pending.clear()
        for group, obj in self._attr_objs.items():
            if group == "":
                self._index["attrs"] = dict(obj)
            else:
                self._index["groups"][group] = dict(obj)
        # Only what changed since the last flush is appended to the journal: the reference flushes after EVERY video
        # (inference_frame_diff.py:299,395,404), and rewriting a whole index of 30 k videos each time would be O(N^2).
        delta = {"groups": {g: self._index["groups"][g] for g in self._dirty_groups if g and g in self._index["groups"]},
                 "datasets": {n: self._index["datasets"][n] for n in self._dirty_ds if n in self._index["datasets"]},
                 "next_file": self._index["next_file"]}
        if "" in self._dirty_groups:
            delta["attrs"] = self._index["attrs"]
        with open(os.path.join(self.path, _JOURNAL), "a", encoding="utf-8") as f:
            f.write(json.dumps(delta) + "\n")
        self._journal_lines += 1
        self._dirty_groups.clear()
        self._dirty_ds.clear()
        self._is_dirty = False

    def close(self):
        if not self._closed:
            self.flush()
            if self.mode != "r" and self._journal_lines:
                self._snapshot()  # compact: one index.json, no journal
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- HDF5 interchange: the exact layout of extract_embeddings.py:50-55,106-119 --
    def _tree(self):
        """The store as the nested structure vimoclip_b200.hdf5_min.write_hdf5 takes."""
        root = {"attrs": dict(self._index["attrs"]), "children": {}}

        def node_of(path):
            cur = root
            for part in path.split("/"):
                cur = cur["children"].setdefault(part, {"attrs": {}, "children": {}})
            return cur

        for g, attrs in self._index["groups"].items():
            if g:
                node_of(g)["attrs"] = dict(attrs)
        for name, meta in self._index["datasets"].items():
            parent, _, leaf = name.rpartition("/")
            holder = node_of(parent) if parent else root
            ds = self._dataset(name)
            if meta["kind"] == "str":
                holder["children"][leaf] = ("dataset", [str(x) for x in ds[:]], {})
            else:
                storage = {"chunks": tuple(meta["chunks"]) if meta.get("chunks") else None, "compression": meta.get("compression")}
                holder["children"][leaf] = ("dataset", np.asarray(ds[...]), {}, storage)
        return root

    def to_hdf5(self, hdf5_path):
        """Write the reference's HDF5 file: one group per video with ``embeddings`` (chunked + gzip as recorded), ``labels``,
        attributes, root attributes, ``video_ids`` as variable-length strings.  With h5py installed it does the writing;
        without it (this image) the built-in writer of ``hdf5_min`` produces the same layout."""
        self.flush()
        try:
            import h5py
        except ImportError:
            from .hdf5_min import write_hdf5

            write_hdf5(hdf5_path, self._tree())
            return
        with h5py.File(hdf5_path, "w") as hf:  # pragma: no cover - h5py is absent from the build image
            for k, v in self._index["attrs"].items():
                hf.attrs[k] = v
            for g, attrs in self._index["groups"].items():
                if g:
                    grp = hf.require_group(g)
                    for k, v in attrs.items():
                        grp.attrs[k] = v
            for name, meta in self._index["datasets"].items():
                ds = self._dataset(name)
                if meta["kind"] == "str":
                    hf.create_dataset(name, data=np.array(list(ds[:]), dtype=h5py.string_dtype()))
                else:
                    kw = {}
                    if meta.get("compression"):
                        kw["compression"] = meta["compression"]
                    if meta.get("chunks"):
                        kw["chunks"] = tuple(meta["chunks"])
                    if meta.get("maxshape"):
                        kw["maxshape"] = tuple(meta["maxshape"])
                    hf.create_dataset(name, data=ds[...], **kw)

    @classmethod
    def from_hdf5(cls, hdf5_path, path):
        """Import an HDF5 file written by the reference (default ``libver``): h5py when installed, else the built-in reader."""
        try:
            import h5py
        except ImportError:
            from .hdf5_min import read_hdf5

            tree = read_hdf5(hdf5_path)
            with cls(path, "w") as st:
                def put(node, group):
                    for k, v in node["attrs"].items():
                        group.attrs[k] = v
                    for name, child in node["children"].items():
                        if isinstance(child, tuple):
                            data = child[1]
                            if data.dtype.kind == "O":
                                data = [str(x) for x in data.ravel()]
                            elif data.dtype.kind == "S":
                                data = [x.decode("utf-8") for x in data.ravel()]
                            group.create_dataset(name, data=data)
                        else:
                            put(child, group.require_group(name))

                put(tree, st)
            return cls(path, "r")
        with h5py.File(hdf5_path, "r") as hf, cls(path, "w") as st:  # pragma: no cover
            for k, v in hf.attrs.items():
                st.attrs[k] = v

            def visit(name, obj):
                if isinstance(obj, h5py.Group):
                    g = st.require_group(name)
                    for k, v in obj.attrs.items():
                        g.attrs[k] = v
                else:
                    data = obj[...]
                    if data.dtype.kind in "OS":
                        data = [s.decode("utf-8") if isinstance(s, bytes) else str(s) for s in data.ravel()]
                    st.create_dataset(name, data=data, compression=obj.compression, chunks=obj.chunks)

            hf.visititems(visit)
        return cls(path, "r")


def write_video(store: Group, video_id: str, embeddings, labels, total_frames: int, original_frames: int, compression: str = "gzip"):
    """One video exactly as ``extract_embeddings.py:106-111`` / ``inference.py:100-110`` lay it out: group ``video_id`` with
    ``embeddings [T, D]`` float32 (gzip, chunks ``(1, D)``), ``labels [C]`` float32, attrs ``total_frames`` / ``original_frames``.
    ``embeddings`` / ``labels`` may be torch tensors on any device."""
    emb = np.ascontiguousarray(_to_numpy(embeddings), dtype=np.float32)
    if emb.ndim != 2:
        raise ValueError("embeddings must be [T, D]")
    g = store.create_group(video_id)
    g.create_dataset("embeddings", data=emb, compression=compression, chunks=(1, emb.shape[1]))
    g.create_dataset("labels", data=np.ascontiguousarray(_to_numpy(labels), dtype=np.float32))
    g.attrs["total_frames"] = int(total_frames)
    g.attrs["original_frames"] = int(original_frames)
    return g


def _to_numpy(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)
