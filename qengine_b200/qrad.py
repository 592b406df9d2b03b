"""ctypes binding of libqrad_cuda.so (include/qrad_cuda.h + include/qrad_host.h).

This is the Python face of the qrad3 drop-in: ``Scene`` wraps the host half
(LoadBSPFile .. CreateDirectLights, WriteBSPFile) and ``Device`` the CUDA half
(the five RadWorld seams, src/tools/qrad3/qrad3.c:494-536).  There is no CPU
fallback: if the shared library is missing or no CUDA device is usable the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libqrad_cuda.so"

c_f = C.POINTER(C.c_float)
c_i = C.POINTER(C.c_int32)
c_u8 = C.POINTER(C.c_uint8)


class QradError(RuntimeError):
    pass


class BspView(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("numplanes", C.c_int32), ("nodes", C.c_void_p), ("numnodes", C.c_int32),
                ("leafs", C.c_void_p), ("numleafs", C.c_int32), ("visdata", C.c_void_p), ("visdatasize", C.c_int32)]


class PatchesView(C.Structure):
    _fields_ = [("num_patches", C.c_int32), ("origin", c_f), ("normal", c_f), ("area", c_f), ("cluster", c_i),
                ("sky", c_u8), ("reflectivity", c_f), ("baselight", c_f), ("totallight", c_f), ("bounds", c_f),
                ("face", c_i), ("next", c_i), ("numfaces", C.c_int32), ("face_head", c_i)]


class LightsView(C.Structure):
    _fields_ = [("numclusters", C.c_int32), ("light_first", c_i), ("numlights", C.c_int32), ("type", c_i),
                ("intensity", c_f), ("style", c_i), ("origin", c_f), ("color", c_f), ("normal", c_f), ("stopdot", c_f)]


class FacesView(C.Structure):
    _fields_ = [("numfaces", C.c_int32), ("lit", c_u8), ("facenormal", c_f), ("texorg", c_f), ("textoworld", c_f),
                ("exactmins", c_f), ("exactmaxs", c_f), ("texmins", c_i), ("texsize", c_i), ("face_bounds", c_f),
                ("planelink_head", c_i), ("facelinks", c_i), ("minlight", c_f), ("plane_normal", c_f)]


class FinalParams(C.Structure):
    _fields_ = [("ambient", C.c_float), ("maxlight", C.c_float), ("lightscale", C.c_float), ("subdiv", C.c_float),
                ("numbounce", C.c_int32)]


class HostParams(C.Structure):
    _fields_ = [("numbounce", C.c_int32), ("extrasamples", C.c_int32), ("subdiv", C.c_float), ("ambient", C.c_float),
                ("maxlight", C.c_float), ("lightscale", C.c_float), ("direct_scale", C.c_float),
                ("entity_scale", C.c_float), ("nopvs", C.c_int32), ("verbose", C.c_int32), ("max_patches", C.c_int32),
                ("threads", C.c_int32)]


class HostCli(C.Structure):
    _fields_ = [("next_arg", C.c_int32), ("threads", C.c_int32), ("tmpin", C.c_int32), ("tmpout", C.c_int32),
                ("dump", C.c_int32), ("glview", C.c_int32), ("device", C.c_int32), ("ngpus", C.c_int32),
                ("direct_given", C.c_int32), ("entity_given", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("rays_transfers", C.c_int64), ("rays_direct", C.c_int64), ("node_visits", C.c_int64),
                ("candidate_pairs", C.c_int64), ("total_transfers", C.c_int64), ("transfer_bytes", C.c_int64),
                ("padded_entries", C.c_int64), ("nonzero_words", C.c_int64),
                ("num_patches", C.c_int32), ("num_luxels", C.c_int32), ("num_clusters", C.c_int32),
                ("tree_depth", C.c_int32),
                ("ms_upload", C.c_float), ("ms_facelights", C.c_float), ("ms_transfers", C.c_float),
                ("ms_transfers_trace", C.c_float), ("ms_bounce", C.c_float), ("ms_bounce_kernel", C.c_float),
                ("ms_final", C.c_float), ("kernel_launches", C.c_int32),
                ("ms_tr_setup", C.c_float), ("ms_tr_lists", C.c_float), ("ms_tr_rows", C.c_float),
                ("ms_tr_columns", C.c_float), ("full_walks", C.c_int64), ("box_pruned_rays", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("ms_x_matrix", C.c_float), ("ms_x_small", C.c_float),
                ("ms_x_radiosity", C.c_float), ("world", C.c_int32), ("rank", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Load libqrad_cuda.so (built in-tree by __graft_entry__.build()). Raises if it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise QradError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.qrad_cuda_last_error.restype = C.c_char_p
        L.qrad_host_last_error.restype = C.c_char_p
        L.qrad_host_free.restype = None
        L.qrad_cuda_destroy.restype = None
        L.qrad_host_default_params.restype = None
        _lib = L
    return _lib


def _hck(rc):
    if rc:
        raise QradError(lib().qrad_host_last_error().decode(errors="replace"))


def _dck(rc):
    if rc:
        raise QradError(lib().qrad_cuda_last_error().decode(errors="replace"))


def _fp(a):
    return a.ctypes.data_as(c_f)


def _ip(a):
    return a.ctypes.data_as(c_i)


def default_params(**kw) -> HostParams:
    p = HostParams()
    lib().qrad_host_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def params_from_flags(flags, with_cli: bool = False):
    """The reference's flag loop (qrad3.c:555-601) as implemented by the library (qrad_host_parse_flags)."""
    p = default_params()
    cli = HostCli()
    args = [b"qrad3"] + [str(a).encode() for a in flags]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    lib().qrad_host_parse_flags(C.c_int(len(args)), arr, C.byref(p), C.byref(cli))
    if cli.next_arg != len(args):
        raise ValueError(f"unknown qrad3 flag {flags[cli.next_arg - 1]!r}")
    return (p, cli) if with_cli else p


def _np_from(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class Scene:
    """Host half: load a .bsp (path must contain .../assets/<dir>/), build patches and lights."""

    def __init__(self, bsp_path):
        self._h = C.c_void_p()
        self.path = str(bsp_path)
        _hck(lib().qrad_host_load(self.path.encode(), C.byref(self._h)))
        self.params = None

    def close(self):
        if self._h:
            lib().qrad_host_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare(self, params: HostParams):
        _hck(lib().qrad_host_prepare(self._h, C.byref(params)))
        self.params = params
        return self

    def prepare_part(self, params: HostParams, part: int, parts: int, copy: bool = True) -> np.ndarray:
        """This process's share of the dicing (qrad_host_prepare_part): a byte blob for prepare_merge of every process.
        copy=False returns the scene's own buffer, valid until the next prepare call on this scene."""
        self.params = params
        blob, n = C.POINTER(C.c_uint8)(), C.c_int64()
        _hck(lib().qrad_host_prepare_part(self._h, C.byref(params), C.c_int(part), C.c_int(parts), C.byref(blob), C.byref(n)))
        a = np.ctypeslib.as_array(blob, shape=(n.value,))
        return a.copy() if copy else a

    def prepare_merge(self, blobs):
        """All processes' blobs, in part order (the own one included): finishes what prepare() does."""
        arrs = [np.ascontiguousarray(b, np.uint8) for b in blobs]
        ptrs = (C.POINTER(C.c_uint8) * len(arrs))(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in arrs])
        sizes = (C.c_int64 * len(arrs))(*[a.size for a in arrs])
        _hck(lib().qrad_host_prepare_merge(self._h, C.c_int(len(arrs)), ptrs, sizes))
        return self

    def counts(self):
        a = (C.c_int32 * 8)()
        lib().qrad_host_counts(self._h, a)
        keys = ["numfaces", "num_patches", "numdlights", "numclusters", "numnodes", "numleafs", "numtexinfo", "num_entities"]
        return dict(zip(keys, list(a)))

    def reflectivity(self):
        n = self.counts()["numtexinfo"]
        out = np.zeros((n, 3), np.float32)
        lib().qrad_host_reflectivity(self._h, _fp(out))
        return out

    def views(self):
        bv, pv, lv, fv = BspView(), PatchesView(), LightsView(), FacesView()
        lib().qrad_host_bsp_view(self._h, C.byref(bv))
        _hck(lib().qrad_host_patches_view(self._h, C.byref(pv)))
        _hck(lib().qrad_host_lights_view(self._h, C.byref(lv)))
        _hck(lib().qrad_host_faces_view(self._h, C.byref(fv)))
        return bv, pv, lv, fv

    def patches(self):
        """Patch arrays as numpy copies."""
        _, pv, _, _ = self.views()
        n = pv.num_patches
        d = dict(
            origin=_np_from(pv.origin, (n, 3), np.float32), normal=_np_from(pv.normal, (n, 3), np.float32),
            area=_np_from(pv.area, (n,), np.float32), cluster=_np_from(pv.cluster, (n,), np.int32),
            sky=_np_from(pv.sky, (n,), np.uint8), reflectivity=_np_from(pv.reflectivity, (n, 3), np.float32),
            baselight=_np_from(pv.baselight, (n, 3), np.float32), totallight=_np_from(pv.totallight, (n, 3), np.float32),
            bounds=_np_from(pv.bounds, (n, 6), np.float32), face=_np_from(pv.face, (n,), np.int32),
            next=_np_from(pv.next, (n,), np.int32), face_head=_np_from(pv.face_head, (pv.numfaces,), np.int32))
        return d

    def lights(self):
        _, _, lv, _ = self.views()
        n = lv.numlights
        return dict(
            light_first=_np_from(lv.light_first, (lv.numclusters + 1,), np.int32),
            type=_np_from(lv.type, (n,), np.int32), intensity=_np_from(lv.intensity, (n,), np.float32),
            style=_np_from(lv.style, (n,), np.int32), origin=_np_from(lv.origin, (n, 3), np.float32),
            color=_np_from(lv.color, (n, 3), np.float32), normal=_np_from(lv.normal, (n, 3), np.float32),
            stopdot=_np_from(lv.stopdot, (n,), np.float32))

    def windings(self):
        n = self.counts()["num_patches"]
        cnt = np.zeros(n, np.int32)
        lib().qrad_host_windings(self._h, _ip(cnt), None)
        pts = np.zeros((int(cnt.sum()), 3), np.float32)
        lib().qrad_host_windings(self._h, _ip(cnt), _fp(pts))
        return cnt, pts

    def lighting(self) -> bytes:
        p = C.POINTER(C.c_uint8)()
        n = C.c_int32()
        lib().qrad_host_lighting(self._h, C.byref(p), C.byref(n))
        return C.string_at(p, n.value)

    def store_lighting(self, data: np.ndarray, lightofs: np.ndarray, styles: np.ndarray):
        _hck(lib().qrad_host_store_lighting(self._h, data.ctypes.data_as(c_u8), C.c_int32(data.size), _ip(lightofs),
                                             styles.ctypes.data_as(c_u8)))

    def write(self, path=None):
        _hck(lib().qrad_host_write(self._h, str(path or self.path).encode()))


class Device:
    """CUDA half: one context on one GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _dck(lib().qrad_cuda_create(C.c_int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().qrad_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_count_nodes(self, on=True):
        _dck(lib().qrad_cuda_set_count_nodes(self._h, C.c_int(1 if on else 0)))

    def upload(self, scene: Scene):
        bv, pv, lv, fv = scene.views()
        _dck(lib().qrad_cuda_upload_bsp(self._h, C.byref(bv)))
        _dck(lib().qrad_cuda_upload_patches(self._h, C.byref(pv)))
        _dck(lib().qrad_cuda_upload_lights(self._h, C.byref(lv)))
        _dck(lib().qrad_cuda_upload_faces(self._h, C.byref(fv)))
        self._n = pv.num_patches
        self._nf = fv.numfaces
        return self

    def upload_bsp(self, scene: Scene):
        bv, _, _, _ = scene.views() if scene.params is not None else (self._bsp_only(scene), None, None, None)
        _dck(lib().qrad_cuda_upload_bsp(self._h, C.byref(bv)))

    @staticmethod
    def _bsp_only(scene):
        bv = BspView()
        lib().qrad_host_bsp_view(scene._h, C.byref(bv))
        return bv

    def test_lines(self, starts: np.ndarray, stops: np.ndarray) -> np.ndarray:
        starts = np.ascontiguousarray(starts, np.float32)
        stops = np.ascontiguousarray(stops, np.float32)
        n = starts.shape[0]
        out = np.zeros(n, np.uint8)
        _dck(lib().qrad_cuda_test_lines(self._h, C.c_int64(n), _fp(starts), _fp(stops), out.ctypes.data_as(c_u8)))
        return out

    def point_in_leaf(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(pts.shape[0], np.int32)
        _dck(lib().qrad_cuda_point_in_leaf(self._h, C.c_int64(pts.shape[0]), _fp(pts), _ip(out)))
        return out

    def build_facelights(self, extrasamples: int, numbounce: int):
        _dck(lib().qrad_cuda_build_facelights(self._h, C.c_int(extrasamples), C.c_int(numbounce)))

    def make_transfers(self, nopvs: int = 0) -> int:
        t = C.c_int64()
        _dck(lib().qrad_cuda_make_transfers(self._h, C.c_int(nopvs), C.byref(t)))
        return t.value

    def bounce(self, numbounce: int, want_added=True):
        added = np.zeros(max(numbounce, 1), np.float32)
        _dck(lib().qrad_cuda_bounce(self._h, C.c_int(numbounce), _fp(added) if want_added else None))
        return added[:numbounce]

    def final_light(self, ambient, maxlight, lightscale, subdiv, numbounce):
        fp = FinalParams(ambient, maxlight, lightscale, subdiv, numbounce)
        cap = 0x200000
        data = np.zeros(cap, np.uint8)
        size = C.c_int32()
        lightofs = np.zeros(self._nf, np.int32)
        styles = np.zeros((self._nf, 4), np.uint8)
        _dck(lib().qrad_cuda_final_light(self._h, C.byref(fp), data.ctypes.data_as(c_u8), C.c_int32(cap), C.byref(size),
                                         _ip(lightofs), styles.ctypes.data_as(c_u8)))
        return data[:size.value].copy(), lightofs, styles

    def stats(self) -> dict:
        s = Stats()
        _dck(lib().qrad_cuda_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def numtransfers(self) -> np.ndarray:
        out = np.zeros(self._n, np.int32)
        _dck(lib().qrad_cuda_download_numtransfers(self._h, _ip(out)))
        return out

    def transfers(self):
        cnt = self.numtransfers()
        nnz = int(cnt.sum())
        row_ptr = np.zeros(self._n + 1, np.int64)
        patch = np.zeros(nnz, np.uint32)
        transfer = np.zeros(nnz, np.uint16)
        _dck(lib().qrad_cuda_download_transfers(self._h, row_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                                patch.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                transfer.ctypes.data_as(C.POINTER(C.c_uint16))))
        return row_ptr, patch, transfer

    def rows(self, patches):
        """Transfer rows of the selected source patches: (row_ptr[n+1], patch u32[], transfer u16[], rays i64[n])."""
        sel = np.ascontiguousarray(patches, np.int32)
        cnt = self.numtransfers()[sel]
        nnz = int(cnt.astype(np.int64).sum())
        row_ptr = np.zeros(sel.size + 1, np.int64)
        patch = np.zeros(max(nnz, 1), np.uint32)
        transfer = np.zeros(max(nnz, 1), np.uint16)
        rays = np.zeros(sel.size, np.int64)
        _dck(lib().qrad_cuda_download_rows(self._h, C.c_int32(sel.size), _ip(sel), row_ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                           patch.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           transfer.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_int64(nnz),
                                           rays.ctypes.data_as(C.POINTER(C.c_int64))))
        return row_ptr, patch[:nnz], transfer[:nnz], rays

    def patch_light(self):
        n = self._n
        tl = np.zeros((n, 3), np.float32)
        sl = np.zeros((n, 3), np.float32)
        sm = np.zeros(n, np.int32)
        rad = np.zeros((n, 3), np.float32)
        _dck(lib().qrad_cuda_download_patch_light(self._h, _fp(tl), _fp(sl), _ip(sm), _fp(rad)))
        return dict(totallight=tl, samplelight=sl, samples=sm, radiosity=rad)

    def upload_patch_light(self, totallight, samplelight, samples):
        tl = np.ascontiguousarray(totallight, np.float32)
        sl = np.ascontiguousarray(samplelight, np.float32)
        sm = np.ascontiguousarray(samples, np.int32)
        _dck(lib().qrad_cuda_upload_patch_light(self._h, _fp(tl), _fp(sl), _ip(sm)))

    def facelight(self):
        """-> luxel_first[F+1], numstyles[F], stylenums[F,32], origins[L,3], samples(list per style slot of [L,3])"""
        nf = self._nf
        first = np.zeros(nf + 1, np.int32)
        nst = np.zeros(nf, np.int32)
        sty = np.zeros((nf, 32), np.int32)
        _dck(lib().qrad_cuda_facelight_layout(self._h, _ip(first), _ip(nst), _ip(sty)))
        L = int(first[-1])
        origins = np.zeros((L, 3), np.float32)
        slots = []
        ids = (C.c_int32 * 256)()
        nids = C.c_int32()
        _dck(lib().qrad_cuda_style_ids(self._h, ids, C.byref(nids)))
        self.style_ids = list(ids)[:nids.value]
        for s in range(max(nids.value, 1)):
            smp = np.zeros((L, 3), np.float32)
            _dck(lib().qrad_cuda_download_facelight(self._h, _fp(origins), C.c_int32(s), _fp(smp)))
            slots.append(smp)
        return first, nst, sty, origins, slots


def rad_world(scene: Scene, dev: Device, params: HostParams, want_added: bool = True):
    """RadWorld (qrad3.c:494-536) through the C entry point; returns `added` per bounce.

    want_added=False skips CollectLight's serial energy sum (only printed by `-v` in the reference)."""
    added = np.zeros(max(params.numbounce, 1), np.float32)
    _hck(lib().qrad_rad_world(scene._h, dev._h, C.byref(params), _fp(added) if want_added else None))
    dev._n = scene.counts()["num_patches"]
    dev._nf = scene.counts()["numfaces"]
    return added[:max(params.numbounce, 0)]


class HostTree:
    """Host builds of the device walkers (include/qrad_cuda.h, qrad_host_tree_*): TestLine, PointInLeaf, the common start
    node and the cell proofs compiled for the CPU from the same source as the kernels' -- for tests, not a fallback."""

    def __init__(self, scene: Scene):
        L = lib()
        L.qrad_host_tree_last_error.restype = C.c_char_p
        L.qrad_host_tree_destroy.restype = None
        bv = BspView()
        L.qrad_host_bsp_view(scene._h, C.byref(bv))
        self._scene = scene          # the view points into the scene's lumps
        self._h = C.c_void_p()
        if L.qrad_host_tree_create(C.byref(bv), C.byref(self._h)):
            raise QradError(L.qrad_host_tree_last_error().decode(errors="replace"))
        info = (C.c_int32 * 4)()
        ext = C.c_float()
        L.qrad_host_tree_info(self._h, info, C.byref(ext))
        self.numnodes, self.numleafs, self.depth, self.celled_solid_leafs = list(info)
        self.extent = ext.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qrad_host_tree_destroy(self._h)
            self._h = None

    def test_lines(self, starts, stops, start_nodes=None):
        starts = np.ascontiguousarray(starts, np.float32)
        stops = np.ascontiguousarray(stops, np.float32)
        n = len(starts)
        res = np.zeros(n, np.uint8)
        blk = np.zeros(n, np.int32)
        sn = None if start_nodes is None else _ip(np.ascontiguousarray(start_nodes, np.int32))
        lib().qrad_host_test_lines(self._h, C.c_int64(n), _fp(starts), _fp(stops), sn, res.ctypes.data_as(c_u8), _ip(blk))
        return res, blk

    def point_in_leaf(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(len(pts), np.int32)
        lib().qrad_host_point_in_leaf(self._h, C.c_int64(len(pts)), _fp(pts), _ip(out))
        return out

    def common_start_node(self, amn, amx, bmn, bmx):
        a = [np.ascontiguousarray(x, np.float32) for x in (amn, amx, bmn, bmx)]
        return lib().qrad_host_common_start_node(self._h, *[_fp(x) for x in a])

    def prove_segments(self, starts, stops, margin=0.0, leaf_lo=0, leaf_hi=1 << 30):
        starts = np.ascontiguousarray(starts, np.float32)
        stops = np.ascontiguousarray(stops, np.float32)
        out = np.zeros(len(starts), np.int32)
        lib().qrad_host_prove_segments(self._h, C.c_float(margin), C.c_int64(len(starts)), _fp(starts), _fp(stops),
                                       C.c_int32(leaf_lo), C.c_int32(min(leaf_hi, self.numleafs)), _ip(out))
        return out

    def prove_boxes(self, amn, amx, bmn, bmx, margin=0.0):
        a = [np.ascontiguousarray(x, np.float32) for x in (amn, amx, bmn, bmx)]
        return lib().qrad_host_prove_boxes(self._h, C.c_float(margin), *[_fp(x) for x in a])


def join_communicator(dev: Device, rank: int, world: int):
    """Make `dev` rank `rank` of a `world`-GPU communicator (one process per GPU, launched by torchrun): rank 0 creates the
    NCCL id inside the library and torch.distributed only carries its 128 bytes to the other ranks.  Afterwards the same
    calls (Device.upload / build_facelights / make_transfers / bounce / final_light, or rad_world) on every rank light the
    map together; every exchange of patch, matrix or radiosity data happens inside libqrad_cuda."""
    import torch
    import torch.distributed as dist
    ident = np.zeros(128, np.uint8)
    if rank == 0:
        _dck(lib().qrad_cuda_comm_unique_id(ident.ctypes.data_as(c_u8)))
    t = torch.from_numpy(ident).cuda()
    dist.broadcast(t, 0)
    ident = t.cpu().numpy()
    _dck(lib().qrad_cuda_comm_init(dev._h, C.c_int(rank), C.c_int(world), ident.ctypes.data_as(c_u8)))


def rad_world_multi(scene: Scene, params: HostParams, ngpus: int, first_device: int = 0, want_added: bool = True):
    """RadWorld on `ngpus` GPUs of this box from ONE process (one host thread per GPU inside the library)."""
    added = np.zeros(max(params.numbounce, 1), np.float32)
    st = Stats()
    _hck(lib().qrad_rad_world_multi(scene._h, C.c_int(first_device), C.c_int(ngpus), C.byref(params),
                                    _fp(added) if want_added else None, C.byref(st)))
    return added[:max(params.numbounce, 0)], st.as_dict()


def qrad3(argv) -> int:
    """The drop-in command line (same flags as the reference qrad3), in-process."""
    args = [b"qrad3"] + [str(a).encode() for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    return lib().qrad3_main(C.c_int(len(args)), arr)
