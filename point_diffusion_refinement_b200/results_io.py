"""Result / resume formats of the reference drivers.

* generated clouds: ``<dataset>_generated_data_<N>pts[_T<t>].h5`` with ONE dataset ``data`` (fp32, (n,N,3), units of
  the un-scaled dataset, i.e. network output / 2 / scale) -- completion_eval.py:281-318.
* per-rank / gathered metrics: a pickled dict with the keys ``meta, cd_distance, emd_distance, f1, avg_cd, avg_emd,
  iter`` -- generate_samples.py:247-252, generate_samples_distributed.py:84-93.

h5py is not part of this image.  When it is importable the .h5 file is written with it, exactly like the reference does;
otherwise ``save_generated`` writes the same container itself (``write_hdf5_dataset``: a minimal HDF5 file in the classic
"libver earliest" layout -- version-0 superblock, symbol-table root group, one contiguous little-endian IEEE fp32
dataset -- following the HDF5 File Format Specification 1.1 byte for byte) and ``load_generated`` reads it back with the
matching minimal reader.  NOTE: there is no libhdf5 in this image, so that writer is validated against the specification
and its own reader only (tests/test_host_model.py has an h5py cross-check that runs wherever h5py exists); whenever the
built-in writer is used a plain ``.npy`` copy is written next to the file (``PDR_RESULTS_NPY=0`` turns it off).
"""
import os
import pickle
import struct

import numpy as np

GENERATED_PREFIX = {"mvp_dataset": "mvp", "shapenet_chunk": "shapenet", "mvp40": "mvp40", "partnet": "partnet"}


def generated_file_name(dataset, num_points, t_slice=None):
    """completion_eval.py:283-290 / :308-315."""
    stem = "%s_generated_data_%dpts" % (GENERATED_PREFIX[dataset], num_points)
    return stem + (".h5" if t_slice is None else "_T%d.h5" % t_slice)


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


_UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages):
    data = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(data)) + data       # version 1 prefix (16 bytes)


def write_hdf5_dataset(path, data, name="data"):
    """One fp32 dataset `name` at the root of a new HDF5 file (format spec III.A-C, IV.A.1-2: superblock v0, object
    header v1, B-tree v1 group node, local heap, symbol table node, contiguous layout v3)."""
    data = np.ascontiguousarray(np.asarray(data, dtype="<f4"))
    nm = name.encode() + b"\0"
    K_LEAF, K_INT = 4, 16
    # ---- addresses -------------------------------------------------------------------------------------------------
    root_oh = 96
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0))])        # placeholder to learn its size
    btree = root_oh + len(root_hdr)
    btree_size = 24 + (2 * K_INT + 1) * 8 + 2 * K_INT * 8
    heap = btree + btree_size
    heap_data = heap + 32
    name_off = 8                                                                   # offset 0 holds the empty root name
    used = _pad8(b"\0") + _pad8(nm)
    heap_seg = used + struct.pack("<QQ", 1, 16)                                    # one free block: next = 1 (last), size 16
    snod = heap_data + len(heap_seg)
    snod_size = 8 + 2 * K_LEAF * 40
    ds_oh = snod + snod_size
    dataspace = struct.pack("<BBBx4x", 1, data.ndim, 0) + b"".join(struct.pack("<Q", d) for d in data.shape)
    datatype = (struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) +                 # class 1 (float) v1, LE, implied msb, sign bit 31
                struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127))                 # offset, precision, exp loc/size, mant loc/size, bias
    fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)                                    # v2: late alloc, write if set, default value
    def ds_header(addr):
        layout = struct.pack("<BBQQ", 3, 1, addr, data.nbytes)                     # v3 contiguous
        return _object_header([_message(0x0001, dataspace), _message(0x0003, datatype, flags=1), _message(0x0005, fill),
                               _message(0x0008, layout)])
    raw = ds_oh + len(ds_header(0))
    raw += -raw % 8
    eof = raw + data.nbytes
    # ---- blocks ----------------------------------------------------------------------------------------------------
    sb = (_SIG + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, K_LEAF, K_INT, 0) +
          struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF) +
          struct.pack("<QQII", 0, root_oh, 1, 0) + struct.pack("<QQ", btree, heap))   # root symbol table entry, cached
    assert len(sb) == 96
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", btree, heap))])
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF) + struct.pack("<QQQ", 0, snod, name_off)
    bt += b"\0" * (btree_size - len(bt))
    hp = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_seg), len(used), heap_data)
    sn = b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", name_off, ds_oh, 0, 0)
    sn += b"\0" * (snod_size - len(sn))
    dh = ds_header(raw)
    with open(path, "wb") as f:
        f.write(sb + root_hdr + bt + hp + heap_seg + sn + dh)
        f.write(b"\0" * (raw - f.tell()))
        f.write(data.tobytes())
    return path


def read_hdf5_dataset(path, name="data"):
    """Reader for the files `write_hdf5_dataset` (and h5py with default settings for a small contiguous fp32 dataset)
    produces: superblock v0 -> root symbol table -> group B-tree -> symbol node -> object header v1 -> contiguous data."""
    buf = open(path, "rb").read()
    assert buf[:8] == _SIG and buf[8] == 0 and buf[13] == 8 and buf[14] == 8, "not a version-0 HDF5 file with 8-byte offsets"
    btree, heap = struct.unpack_from("<QQ", buf, 56 + 24)
    assert buf[heap:heap + 4] == b"HEAP"
    heap_data = struct.unpack_from("<Q", buf, heap + 24)[0]

    def heap_name(off):
        end = buf.index(b"\0", heap_data + off)
        return buf[heap_data + off:end].decode()

    def leaves(node):
        assert buf[node:node + 4] == b"TREE"
        _, level, used = struct.unpack_from("<BBH", buf, node + 4)
        for i in range(used):
            child = struct.unpack_from("<Q", buf, node + 24 + 8 + 16 * i)[0]
            if level == 0:
                yield child
            else:
                yield from leaves(child)

    target = None
    for sn in leaves(btree):
        assert buf[sn:sn + 4] == b"SNOD"
        for i in range(struct.unpack_from("<H", buf, sn + 6)[0]):
            off, oh = struct.unpack_from("<QQ", buf, sn + 8 + 40 * i)
            if heap_name(off) == name:
                target = oh
    if target is None:
        raise KeyError(name)
    version, nmsg, _, hsize = struct.unpack_from("<BxHII", buf, target)
    assert version == 1
    pos, end = target + 16, target + 16 + hsize
    shape = dtype = addr = None
    while pos < end and nmsg > 0:
        mtype, msize = struct.unpack_from("<HH", buf, pos)
        body = pos + 8
        if mtype == 0x0001:
            rank = buf[body + 1]
            shape = struct.unpack_from("<%dQ" % rank, buf, body + 8)
        elif mtype == 0x0003:
            cls, size = buf[body] & 0x0F, struct.unpack_from("<I", buf, body + 4)[0]
            assert cls == 1 and size in (4, 8) and not (buf[body + 1] & 1), "little-endian IEEE float datasets only"
            dtype = "<f%d" % size
        elif mtype == 0x0008:
            assert buf[body] == 3 and buf[body + 1] == 1, "contiguous layout only"
            addr = struct.unpack_from("<Q", buf, body + 2)[0]
        pos = body + msize
        nmsg -= 1
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape)), offset=addr).reshape(shape).copy()


def save_generated(path, data):
    """Write `data` (n,N,3) as dataset 'data' of the HDF5 file `path`; returns the file actually written."""
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float32))
    h5 = _h5py()
    if h5 is not None:
        with h5.File(path, "w") as hf:
            hf.create_dataset("data", data=data)
    else:
        write_hdf5_dataset(path, data, "data")
    # the built-in writer cannot be checked against libhdf5 in this image: unless h5py wrote the file, a plain .npy copy
    # goes next to it (PDR_RESULTS_NPY=0 turns that off, =1 forces it)
    npy = os.environ.get("PDR_RESULTS_NPY")
    if npy == "1" or (npy != "0" and h5 is None):
        np.save(os.path.splitext(path)[0] + ".npy", data)
    return path


def load_generated(path):
    h5 = _h5py()
    if os.path.exists(path):
        if h5 is not None:
            with h5.File(path, "r") as hf:
                return np.array(hf["data"])
        return read_hdf5_dataset(path, "data")
    alt = os.path.splitext(path)[0] + ".npy"
    if os.path.exists(alt):
        return np.load(alt)
    raise FileNotFoundError("%s / %s" % (path, alt))


def eval_result_dict(meta, cd_distance, emd_distance, f1, iteration):
    """The dict the reference pickles per rank and after gathering (generate_samples.py:247-252)."""
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    cd, emd = to_np(cd_distance), to_np(emd_distance)
    return {"meta": to_np(meta), "cd_distance": cd, "emd_distance": emd, "f1": to_np(f1),
            "avg_cd": cd.mean(), "avg_emd": emd.mean(), "iter": iteration}


def save_eval_result(path, result):
    with open(path, "wb") as handle:
        pickle.dump(result, handle)
    return path


def load_eval_result(path):
    with open(path, "rb") as handle:
        return pickle.load(handle)


def gather_eval_results(results):
    """Concatenate per-rank dicts in rank order (generate_samples_distributed.py:60-93)."""
    cat = lambda k: np.concatenate([np.asarray(r[k]) for r in results], axis=0)
    cd, emd = cat("cd_distance"), cat("emd_distance")
    return {"meta": cat("meta"), "cd_distance": cd, "emd_distance": emd, "f1": cat("f1"),
            "avg_cd": cd.mean(), "avg_emd": emd.mean(), "iter": results[-1]["iter"]}
