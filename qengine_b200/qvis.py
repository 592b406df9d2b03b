"""ctypes mirror of include/qvis_cuda.h: the first stage of the reference's qvis3 (LoadPortals, BasePortalVis, SimpleFlood)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import qrad as Q

c_f = C.POINTER(C.c_float)
c_i = C.POINTER(C.c_int32)
c_u8 = C.POINTER(C.c_uint8)


class PortalsView(C.Structure):
    _fields_ = [("numportals", C.c_int32), ("numclusters", C.c_int32), ("planes", c_f), ("leaf", c_i),
                ("winding_first", c_i), ("points", c_f), ("leaf_first", c_i), ("leaf_portals", c_i)]


class QvisError(RuntimeError):
    pass


def _lib():
    lib = Q.lib()
    lib.qvis_last_error.restype = C.c_char_p
    lib.qvis_host_portals_free.restype = None
    return lib


def _ck(rc):
    if rc != 0:
        raise QvisError(_lib().qvis_last_error().decode())


class Portals:
    """LoadPortals (qvis3.c:304-412): the memory portals of a .prt file."""

    def __init__(self, prt_path):
        self._h = C.c_void_p()
        _ck(_lib().qvis_host_portals_load(str(prt_path).encode(), C.byref(self._h)))
        self.view = PortalsView()
        _ck(_lib().qvis_host_portals_view(self._h, C.byref(self.view)))
        v = self.view
        self.numportals, self.numclusters = v.numportals, v.numclusters
        self.portalbytes = _lib().qvis_portal_bytes(C.c_int32(v.numportals))

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().qvis_host_portals_free(self._h)
            self._h = None

    def arrays(self):
        v, P, nc = self.view, self.numportals, self.numclusters
        wf = np.ctypeslib.as_array(v.winding_first, shape=(P + 1,)).copy()
        lf = np.ctypeslib.as_array(v.leaf_first, shape=(nc + 1,)).copy()
        return dict(planes=np.ctypeslib.as_array(v.planes, shape=(P, 4)).copy(), leaf=np.ctypeslib.as_array(v.leaf, shape=(P,)).copy(),
                    winding_first=wf, points=np.ctypeslib.as_array(v.points, shape=(int(wf[-1]), 3)).copy(), leaf_first=lf,
                    leaf_portals=np.ctypeslib.as_array(v.leaf_portals, shape=(int(lf[-1]),)).copy())

    def base_portal_vis(self, device: int = 0):
        """BasePortalVis + SimpleFlood for every portal on the GPU: (portalfront, portalflood) [P, portalbytes] uint8,
        nummightsee [P], kernel times in ms."""
        P, pb = self.numportals, self.portalbytes
        front = np.zeros((P, pb), np.uint8)
        flood = np.zeros((P, pb), np.uint8)
        might = np.zeros(P, np.int32)
        t0, t1 = C.c_float(), C.c_float()
        _ck(_lib().qvis_cuda_base_portal_vis(C.c_int(device), C.byref(self.view), front.ctypes.data_as(c_u8), flood.ctypes.data_as(c_u8),
                                             might.ctypes.data_as(c_i), C.byref(t0), C.byref(t1)))
        return front, flood, might, (t0.value, t1.value)

    def merge_lump(self, portalvis, cap: int = 0x100000):
        """ClusterMerge + CalcPHS + CompressVis on the host: (lump bytes, average clusters visible, hearable, warnings)."""
        pv = np.ascontiguousarray(portalvis, np.uint8)
        assert pv.shape == (self.numportals, self.portalbytes)
        lump = np.zeros(cap, np.uint8)
        size, vis, hear, warned = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _ck(_lib().qvis_host_merge_lump(self._h, pv.ctypes.data_as(c_u8), lump.ctypes.data_as(c_u8), C.c_int32(cap), C.byref(size),
                                        C.byref(vis), C.byref(hear), C.byref(warned)))
        return lump[:size.value].tobytes(), vis.value, hear.value, warned.value


def fast_vis(bsp_in, prt_path, bsp_out, device: int = 0):
    """`qvis3 -fast`: BasePortalVis on the GPU, ClusterMerge / CalcPHS on the host, writes bsp_out."""
    _ck(_lib().qvis_host_fast_vis(str(bsp_in).encode(), str(prt_path).encode(), str(bsp_out).encode(), C.c_int(device)))
