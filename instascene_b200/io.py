"""On-disk formats either side of the hot path, numpy only -- no plyfile / open3d:

* ``point_cloud.ply`` of a trained scene, exactly the vertex layout ``GaussianModel.save_ply`` writes and
  ``GaussianModel.load_ply`` reads (scene/gaussian_model.py:263-313, 364-422): binary little-endian float32 properties
  ``x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0 scale_1 rot_0..3 [segfeat_0..F-1]``; the SH blocks are stored
  channel-major (``[P,15,3] -> transpose(1,2) -> flatten``).
* COLMAP ``cameras.bin`` / ``images.bin`` (PINHOLE), the binary layouts ``read_intrinsics_binary`` /
  ``read_extrinsics_binary`` parse (scene/colmap_loader.py:180-241), plus the conversion to R/T/FoV that
  ``readColmapCameras`` applies (scene/dataset_readers.py:68-101).  The writers let the benchmark's synthetic ring of
  views (BASELINE.json configs[3]: "200 synthetic COLMAP views") travel through the same files a real scene uses.
"""
from __future__ import annotations

import math
import os
import struct
from typing import Dict, List, NamedTuple, Optional

import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# point_cloud.ply
# ---------------------------------------------------------------------------------------------------------------------
_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def gaussian_ply_attributes(n_dc: int, n_rest: int, n_scale: int, n_rot: int, n_segfeat: int) -> List[str]:
    """construct_list_of_attributes (scene/gaussian_model.py:263-283), export_as_3dgs=False."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    names += [f"segfeat_{i}" for i in range(n_segfeat)]
    return names


def save_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation, seg_feature=None,
             crop_mask=None) -> None:
    """Arrays in the model's own layout: xyz [P,3], features_dc [P,1,3], features_rest [P,M-1,3], opacity [P,1]
    (pre-activation), scaling [P,2] (log), rotation [P,4], seg_feature [P,F] or None."""
    a = lambda t: np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
    xyz = a(xyz)
    P = xyz.shape[0]
    f_dc = a(features_dc).transpose(0, 2, 1).reshape(P, -1)
    f_rest = a(features_rest).transpose(0, 2, 1).reshape(P, -1)
    cols = [xyz, np.zeros_like(xyz), f_dc, f_rest, a(opacity).reshape(P, 1), a(scaling).reshape(P, -1),
            a(rotation).reshape(P, -1)]
    n_seg = 0
    if seg_feature is not None:
        seg = a(seg_feature).reshape(P, -1)
        n_seg = seg.shape[1]
        cols.append(seg)
    table = np.ascontiguousarray(np.concatenate(cols, axis=1).astype("<f4"))
    if crop_mask is not None:
        table = np.ascontiguousarray(table[np.asarray(crop_mask, dtype=bool)])
    names = gaussian_ply_attributes(f_dc.shape[1], f_rest.shape[1], cols[5].shape[1], cols[6].shape[1], n_seg)
    assert len(names) == table.shape[1]
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(table.tobytes())


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """Reads the first element ('vertex') of an ascii or binary (little/big endian) PLY into {property: array}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_first, seen_elements = None, 0, [], False, 0
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                seen_elements += 1
                in_first = seen_elements == 1
                if in_first:
                    count = int(tok[2])
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties in the vertex element are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=count, ndmin=2, dtype=np.float64) if count else np.zeros((0, len(props)))
            return {n: rows[:, i].astype(t) for i, (n, t) in enumerate(props)}
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt!r}")
        end = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, end + t) for n, t in props])
        data = np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
        return {n: np.ascontiguousarray(data[n]) for n, _ in props}


class GaussianArrays(NamedTuple):
    xyz: np.ndarray            # [P,3]
    features_dc: np.ndarray    # [P,1,3]
    features_rest: np.ndarray  # [P,(D+1)^2-1,3]
    opacity: np.ndarray        # [P,1]
    scaling: np.ndarray        # [P,2]
    rotation: np.ndarray       # [P,4]
    seg_feature: Optional[np.ndarray]  # [P,F] or None


def load_ply(path: str, max_sh_degree: int = 3, seg_feat_dim: Optional[int] = None) -> GaussianArrays:
    """GaussianModel.load_ply (scene/gaussian_model.py:364-418) on numpy arrays, float32, in the model's layout.
    Like the reference: f_rest_* / scale_* / rot_* are ordered by their numeric suffix, only the first two scale
    columns are used, the number of f_rest columns must match the SH degree, and the seg feature is returned only when
    the file holds exactly `seg_feat_dim` segfeat columns (any number when seg_feat_dim is None)."""
    v = read_ply_vertices(path)
    f32 = lambda n: np.asarray(v[n], dtype=np.float32)
    by_suffix = lambda prefix: sorted((n for n in v if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    xyz = np.stack([f32("x"), f32("y"), f32("z")], axis=1)
    P = xyz.shape[0]
    features_dc = np.stack([f32("f_dc_0"), f32("f_dc_1"), f32("f_dc_2")], axis=1).reshape(P, 3, 1)
    rest_names = by_suffix("f_rest_")
    n_rest = (max_sh_degree + 1) ** 2 - 1
    assert len(rest_names) == 3 * n_rest, (len(rest_names), 3 * n_rest)
    rest = np.stack([f32(n) for n in rest_names], axis=1).reshape(P, 3, n_rest) if rest_names else np.zeros((P, 3, 0), np.float32)
    scaling = np.stack([f32(n) for n in by_suffix("scale_")[:2]], axis=1)
    rotation = np.stack([f32(n) for n in by_suffix("rot")], axis=1)
    seg_names = by_suffix("segfeat")
    seg = None
    if seg_names and (seg_feat_dim is None or len(seg_names) == seg_feat_dim):
        seg = np.stack([f32(f"segfeat_{i}") for i in range(len(seg_names))], axis=1)
    return GaussianArrays(xyz, np.ascontiguousarray(features_dc.transpose(0, 2, 1)),
                          np.ascontiguousarray(rest.transpose(0, 2, 1)), f32("opacity").reshape(P, 1), scaling, rotation, seg)


# ---------------------------------------------------------------------------------------------------------------------
# COLMAP binary model (cameras.bin / images.bin)
# ---------------------------------------------------------------------------------------------------------------------
_CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                  5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                  9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}
_CAMERA_MODEL_IDS = {name: (mid, n) for mid, (name, n) in _CAMERA_MODELS.items()}


class ColmapCamera(NamedTuple):
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


class ColmapImage(NamedTuple):
    id: int
    qvec: np.ndarray   # (w, x, y, z), world -> camera
    tvec: np.ndarray
    camera_id: int
    name: str
    xys: np.ndarray
    point3D_ids: np.ndarray


def qvec2rotmat(q) -> np.ndarray:
    w, x, y, z = (float(c) for c in q)
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def rotmat2qvec(R) -> np.ndarray:
    """Largest-eigenvector construction (same convention as COLMAP: w >= 0)."""
    R = np.asarray(R, dtype=np.float64)
    K = np.array([[R[0, 0] - R[1, 1] - R[2, 2], 0, 0, 0],
                  [R[1, 0] + R[0, 1], R[1, 1] - R[0, 0] - R[2, 2], 0, 0],
                  [R[2, 0] + R[0, 2], R[2, 1] + R[1, 2], R[2, 2] - R[0, 0] - R[1, 1], 0],
                  [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


def write_cameras_binary(path: str, cameras: Dict[int, ColmapCamera]) -> None:
    """Layout parsed by read_intrinsics_binary (scene/colmap_loader.py:215-241): u64 count, then per camera
    i32 id, i32 model id, u64 width, u64 height, f64 params[num_params(model)]."""
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for cam in cameras.values():
            mid, n = _CAMERA_MODEL_IDS[cam.model]
            assert len(cam.params) == n, (cam.model, len(cam.params))
            f.write(struct.pack("<iiQQ", cam.id, mid, cam.width, cam.height))
            f.write(struct.pack("<" + "d" * n, *[float(p) for p in cam.params]))


def read_cameras_binary(path: str) -> Dict[int, ColmapCamera]:
    out = {}
    with open(path, "rb") as f:
        (n_cam,) = struct.unpack("<Q", f.read(8))
        for _ in range(n_cam):
            cid, mid, w, h = struct.unpack("<iiQQ", f.read(24))
            name, n = _CAMERA_MODELS[mid]
            params = np.array(struct.unpack("<" + "d" * n, f.read(8 * n)))
            out[cid] = ColmapCamera(cid, name, w, h, params)
    assert len(out) == n_cam
    return out


def write_images_binary(path: str, images: Dict[int, ColmapImage]) -> None:
    """Layout parsed by read_extrinsics_binary (scene/colmap_loader.py:180-212): u64 count, then per image
    i32 id, f64 qvec[4], f64 tvec[3], i32 camera id, NUL-terminated name, u64 n2D, n2D x (f64 x, f64 y, i64 point3D id)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for im in images.values():
            f.write(struct.pack("<idddddddi", im.id, *[float(c) for c in im.qvec], *[float(c) for c in im.tvec], im.camera_id))
            f.write(im.name.encode("utf-8") + b"\x00")
            n2d = 0 if im.xys is None else len(im.xys)
            f.write(struct.pack("<Q", n2d))
            if n2d:
                rec = np.zeros(n2d, dtype=[("x", "<f8"), ("y", "<f8"), ("id", "<i8")])
                rec["x"], rec["y"], rec["id"] = im.xys[:, 0], im.xys[:, 1], im.point3D_ids
                f.write(rec.tobytes())


def read_images_binary(path: str) -> Dict[int, ColmapImage]:
    out = {}
    with open(path, "rb") as f:
        (n_img,) = struct.unpack("<Q", f.read(8))
        for _ in range(n_img):
            vals = struct.unpack("<idddddddi", f.read(64))
            name = bytearray()
            while True:
                ch = f.read(1)
                if ch == b"\x00" or ch == b"":
                    break
                name += ch
            (n2d,) = struct.unpack("<Q", f.read(8))
            rec = np.frombuffer(f.read(24 * n2d), dtype=[("x", "<f8"), ("y", "<f8"), ("id", "<i8")], count=n2d)
            out[vals[0]] = ColmapImage(vals[0], np.array(vals[1:5]), np.array(vals[5:8]), vals[8], name.decode("utf-8"),
                                       np.column_stack([rec["x"], rec["y"]]) if n2d else np.zeros((0, 2)),
                                       rec["id"].astype(np.int64))
    return out


def focal2fov(focal: float, pixels: float) -> float:   # utils/graphics_utils.py:76-77
    return 2.0 * math.atan(pixels / (2.0 * focal))


def fov2focal(fov: float, pixels: float) -> float:     # utils/graphics_utils.py:73-74
    return pixels / (2.0 * math.tan(fov / 2.0))


def write_synthetic_colmap(sparse_dir: str, cams) -> None:
    """Writes a list of synth.SynthCamera (c2w rotation R, w2c translation T, FoVs) as a PINHOLE COLMAP model:
    one shared camera (id 1) when all views share size and FoV, `images.bin` entries named view_00000.png ..."""
    os.makedirs(sparse_dir, exist_ok=True)
    cameras, images, key_to_id = {}, {}, {}
    for i, c in enumerate(cams):
        key = (c.image_width, c.image_height, c.FoVx, c.FoVy)
        if key not in key_to_id:
            cid = len(key_to_id) + 1
            key_to_id[key] = cid
            fx, fy = fov2focal(c.FoVx, c.image_width), fov2focal(c.FoVy, c.image_height)
            cameras[cid] = ColmapCamera(cid, "PINHOLE", c.image_width, c.image_height,
                                        np.array([fx, fy, c.image_width / 2.0, c.image_height / 2.0]))
        # the loader takes R = qvec2rotmat(qvec).T (dataset_readers.py:80): qvec encodes the w2c rotation R^T
        images[i + 1] = ColmapImage(i + 1, rotmat2qvec(np.asarray(c.R).T), np.asarray(c.T, dtype=np.float64), key_to_id[key],
                                    f"view_{i:05d}.png", None, None)
    write_cameras_binary(os.path.join(sparse_dir, "cameras.bin"), cameras)
    write_images_binary(os.path.join(sparse_dir, "images.bin"), images)


def load_colmap_cameras(sparse_dir: str):
    """readColmapCameras (scene/dataset_readers.py:68-101) without the image files: list of synth.SynthCamera built
    from cameras.bin / images.bin exactly as the reference derives R, T, FovX, FovY (sorted by image name,
    dataset_readers.py:162)."""
    from . import synth
    intr = read_cameras_binary(os.path.join(sparse_dir, "cameras.bin"))
    extr = read_images_binary(os.path.join(sparse_dir, "images.bin"))
    cams = []
    for im in sorted(extr.values(), key=lambda x: x.name):
        cam = intr[im.camera_id]
        R = qvec2rotmat(im.qvec).T
        T = np.array(im.tvec)
        if cam.model in ("SIMPLE_PINHOLE", "SIMPLE_RADIAL"):
            fx = fy = cam.params[0]
        elif cam.model in ("PINHOLE", "OPENCV"):
            fx, fy = cam.params[0], cam.params[1]
        else:
            raise ValueError("Colmap camera model not handled: only undistorted datasets (PINHOLE or SIMPLE_PINHOLE or "
                             "SIMPLE_RADIAL cameras) supported!")
        cams.append(synth.make_camera(R, T, int(cam.width), int(cam.height), focal2fov(fx, cam.width),
                                      focal2fov(fy, cam.height)))
    return cams
