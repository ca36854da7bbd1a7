"""Packed on-disk trajectory format.

The reference keeps one pickle per (size factor, sample): ``pickle((teacher_list, student_list))`` of
``(tensor, t)`` tuples (utils/trajectory_manager.py:232-261).  A pack holds MANY samples of one size factor as
two dense arrays -- exactly the layout the samplers write on the device and the metric kernels read -- in one
uncompressed ``.npz``:

    trajectory_size_{sf}_pack_{first}_{count}.npz
        teacher   float32 [N, Lt, C, H, W]      student   float32 [N, Ls, C, H, W]
        teacher_t int64   [Lt]                  student_t int64   [Ls]         (timestep of every stored frame)
        samples   int64   [N]                   (sample indices, the ``i`` of ``..._sample_{i}.pkl``)

``as_reference_lists`` turns a pack back into the reference structure, so ``load_trajectories`` callers cannot
tell the formats apart.  No torch / CUDA dependency in this module.
"""
import os
import re

import numpy as np

_PACK = re.compile(r"^trajectory_size_(?P<sf>.+)_pack_(?P<first>\d+)_(?P<count>\d+)\.npz$")
_PKL = re.compile(r"^trajectory_size_(?P<sf>.+)_sample_(?P<i>\d+)\.pkl$")


def pack_name(size_factor, first, count):
    return f"trajectory_size_{size_factor}_pack_{first}_{count}.npz"


def write_pack(directory, size_factor, samples, teacher, student, teacher_t, student_t):
    """teacher / student: float32 arrays [N, L, C, H, W]; samples: the N sample indices."""
    samples = np.asarray(samples, np.int64)
    teacher = np.ascontiguousarray(teacher, np.float32)
    student = np.ascontiguousarray(student, np.float32)
    if teacher.ndim != 5 or student.ndim != 5 or teacher.shape[0] != len(samples) or student.shape[0] != len(samples):
        raise ValueError("write_pack expects [N, L, C, H, W] arrays and N sample indices")
    if teacher.shape[1] != len(teacher_t) or student.shape[1] != len(student_t):
        raise ValueError("one timestep value per stored frame is required")
    path = os.path.join(directory, pack_name(size_factor, int(samples.min()) if len(samples) else 0, len(samples)))
    tmp = path + ".tmp.npz"
    np.savez(tmp, teacher=teacher, student=student, teacher_t=np.asarray(teacher_t, np.int64),
             student_t=np.asarray(student_t, np.int64), samples=samples)
    os.replace(tmp, path)            # a crash never leaves a half-written pack behind
    return path


def list_packs(directory, size_factor):
    out = []
    for f in os.listdir(directory):
        m = _PACK.match(f)
        if m and m.group("sf") == str(size_factor):
            out.append((int(m.group("first")), os.path.join(directory, f)))
    return [p for _, p in sorted(out)]


def list_pickles(directory, size_factor):
    """(sample index, path) of the reference-format files, sorted by index (trajectory_manager.py:407-411)."""
    out = []
    for f in os.listdir(directory):
        m = _PKL.match(f)
        if m and m.group("sf") == str(size_factor):
            out.append((int(m.group("i")), os.path.join(directory, f)))
    return sorted(out)


def _mmap_member(path, zf, name):
    """Memory-map one STORED (uncompressed) ``.npy`` member of an ``.npz`` archive; None if it cannot be mapped.
    (``np.load(path, mmap_mode="r")`` silently ignores mmap_mode for archives and reads every member into memory.)"""
    import struct
    import zipfile
    try:
        info = zf.getinfo(name + ".npy")
    except KeyError:
        return None
    if info.compress_type != zipfile.ZIP_STORED:
        return None
    with open(path, "rb") as f:
        f.seek(info.header_offset)
        hdr = f.read(30)                                     # local file header: sizes of its name / extra fields
        if hdr[:4] != b"PK\x03\x04":
            return None
        n_name, n_extra = struct.unpack("<HH", hdr[26:30])
        f.seek(info.header_offset + 30 + n_name + n_extra)
        major, _ = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0 if major == 1 else np.lib.format.read_array_header_2_0)(f)
        if fortran or dtype.hasobject:
            return None
        return np.memmap(path, dtype=dtype, mode="r", offset=f.tell(), shape=shape)


def read_pack(path, mmap=True):
    """dict of arrays; ``teacher`` / ``student`` are memory-mapped straight out of the (uncompressed) archive, so
    reading a few samples of a large pack touches only their pages."""
    import zipfile
    out = {}
    with zipfile.ZipFile(path) as zf, np.load(path) as z:
        for k in ("teacher", "student", "teacher_t", "student_t", "samples"):
            a = _mmap_member(path, zf, k) if mmap and k in ("teacher", "student") else None
            out[k] = a if a is not None else z[k]
    return out


def stored_samples(directory, size_factor):
    """All sample indices present in either format."""
    have = {i for i, _ in list_pickles(directory, size_factor)}
    for p in list_packs(directory, size_factor):
        with np.load(p) as z:
            have.update(int(s) for s in z["samples"])
    return have


def as_reference_lists(pack, to_tensor, indices=None):
    """Pack -> (teacher_trajectories, student_trajectories): per sample a list of ``(tensor [1,C,H,W], t)``
    tuples, the structure TrajectoryManager.load_trajectories returns (trajectory_manager.py:389-432).
    ``to_tensor`` converts one [1,C,H,W] float32 array (e.g. ``torch.from_numpy``)."""
    T, S = [], []
    for n, i in enumerate(pack["samples"]):
        if indices is not None and int(i) not in indices:
            continue
        T.append([(to_tensor(np.array(pack["teacher"][n, k][None])), int(t)) for k, t in enumerate(pack["teacher_t"])])
        S.append([(to_tensor(np.array(pack["student"][n, k][None])), int(t)) for k, t in enumerate(pack["student_t"])])
    return T, S
