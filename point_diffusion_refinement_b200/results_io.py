"""Result / resume formats of the reference drivers.

* generated clouds: ``<dataset>_generated_data_<N>pts[_T<t>].h5`` with ONE dataset ``data`` (fp32, (n,N,3), units of
  the un-scaled dataset, i.e. network output / 2 / scale) -- completion_eval.py:281-318.
* per-rank / gathered metrics: a pickled dict with the keys ``meta, cd_distance, emd_distance, f1, avg_cd, avg_emd,
  iter`` -- generate_samples.py:247-252, generate_samples_distributed.py:84-93.

h5py is not part of this image; when it is importable the .h5 file is written exactly like the reference does,
otherwise the same array goes to ``<name>.npy`` next to where the .h5 would be (and ``load_generated`` reads either).
"""
import os
import pickle

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


def save_generated(path, data):
    """Write `data` (n,N,3) as dataset 'data' of the HDF5 file `path`; returns the file actually written."""
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float32))
    h5 = _h5py()
    if h5 is not None:
        with h5.File(path, "w") as hf:
            hf.create_dataset("data", data=data)
        return path
    alt = os.path.splitext(path)[0] + ".npy"
    np.save(alt, data)
    return alt


def load_generated(path):
    h5 = _h5py()
    if os.path.exists(path) and h5 is not None:
        with h5.File(path, "r") as hf:
            return np.array(hf["data"])
    alt = os.path.splitext(path)[0] + ".npy"
    if os.path.exists(alt):
        return np.load(alt)
    raise FileNotFoundError("%s (h5py %s) / %s" % (path, "present" if h5 else "absent", alt))


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
