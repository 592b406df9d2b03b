"""Python glue of the CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by qengine_b200/ or the C ABI).

Parses a .bsp (IBSP v38, src/tools/common/qfiles.h:252-469), its entity string (common/bspfile.c:625-736),
the palette and .wal textures, feeds oracle/liboracle.so (qrad_oracle.c) and runs RadWorld
(src/tools/qrad3/qrad3.c:494-536) phase by phase on the CPU.
"""
from __future__ import annotations

import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
LUMP = dict(entities=0, planes=1, vertexes=2, vis=3, nodes=4, texinfo=5, faces=6, lighting=7, leafs=8, edges=11,
            surfedges=12, models=13)


def build():
    subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)


def _lib():
    if not LIB.exists():
        build()
    L = C.CDLL(str(LIB))
    L.oq_total_transfer.restype = C.c_longlong
    return L


def parse_entities(text: bytes):
    """ParseEntities with the tokenizer rules of common/scriplib.c:157-250; returns a list of dicts (last key wins)."""
    ents, i, n = [], 0, len(text)

    def token():
        nonlocal i
        while True:
            while i < n and text[i] <= 32:
                i += 1
            if i >= n:
                return None
            if text[i:i + 1] in (b";", b"#") or text[i:i + 2] == b"//":
                while i < n and text[i] != 10:
                    i += 1
                continue
            break
        if text[i] == 34:
            j = text.index(b'"', i + 1) if b'"' in text[i + 1:] else n
            t = text[i + 1:j]
            i = j + 1
        else:
            j = i
            while j < n and text[j] > 32 and text[j] != 59:
                j += 1
            t = text[i:j]
            i = j
        return t.split(b"\0")[0]

    while True:
        t = token()
        if t is None:
            break
        assert t == b"{", t
        e = {}
        while True:
            k = token()
            if k == b"}":
                break
            v = token()
            e[k.decode().rstrip(" \t\r\n")] = v.decode().rstrip(" \t\r\n")
        ents.append(e)
    return ents


class Oracle:
    def __init__(self, bsp_path, numbounce=8, extrasamples=0, subdiv=64.0, ambient=0.0, maxlight=196.0, lightscale=1.0,
                 direct_scale=0.4, entity_scale=1.0, nopvs=0):
        self.L = _lib()
        self.path = Path(bsp_path)
        d = self.path.read_bytes()
        ident, version = struct.unpack_from("<ii", d, 0)
        assert ident == 0x50534249 and version == 38
        self.lumps = {}
        for name, idx in LUMP.items():
            ofs, ln = struct.unpack_from("<ii", d, 8 + idx * 8)
            self.lumps[name] = d[ofs:ofs + ln]
        lu = self.lumps
        self.numfaces = len(lu["faces"]) // 20
        self.numtexinfo = len(lu["texinfo"]) // 76
        self.nummodels = len(lu["models"]) // 48
        self.L.oq_set_bsp(lu["planes"], len(lu["planes"]) // 20, lu["nodes"], len(lu["nodes"]) // 28, lu["leafs"],
                          len(lu["leafs"]) // 28, lu["vis"], len(lu["vis"]), lu["faces"], self.numfaces, lu["texinfo"],
                          self.numtexinfo, lu["vertexes"], len(lu["vertexes"]) // 12, lu["edges"], len(lu["edges"]) // 4,
                          lu["surfedges"], len(lu["surfedges"]) // 4, lu["models"], self.nummodels)
        self.L.oq_set_params(numbounce, extrasamples, C.c_float(subdiv), C.c_float(ambient), C.c_float(maxlight),
                             C.c_float(lightscale), C.c_float(direct_scale), C.c_float(entity_scale), nopvs)
        self.numbounce = numbounce if len(lu["vis"]) else 0
        self.entities = parse_entities(lu["entities"])
        self._reflectivity()
        self._model_entities()

    # CalcTextureReflectivity, patches.c:40-112 (file handling here, arithmetic in C)
    def _reflectivity(self):
        gamedir = self.path.parent      # .../assets/<dir>/
        pcx = (gamedir / ".." / "pics" / "colormap.pcx").read_bytes()
        palette = pcx[-768:]
        names = []
        for i in range(self.numtexinfo):
            name = self.lumps["texinfo"][i * 76 + 40:i * 76 + 72].split(b"\0")[0].decode()
            if name in names:
                self.L.oq_copy_reflectivity(i, names.index(name))
            else:
                wal = gamedir / ".." / "textures" / (name + ".wal")
                if wal.exists():
                    w = wal.read_bytes()
                    width, height, ofs0 = struct.unpack_from("<III", w, 32)
                    self.L.oq_texture_reflectivity(i, palette, w[ofs0:ofs0 + width * height], width * height)
                else:
                    self.L.oq_texture_reflectivity(i, palette, None, 0)
            names.append(name)

    def _model_entities(self):  # EntityForModel, patches.c:253-268
        for m in range(self.nummodels):
            ent = next((e for e in self.entities if e.get("model", "") == "*%d" % m), self.entities[0] if self.entities else {})
            self.L.oq_set_model_entity(m, ent.get("origin", "").encode(), ent.get("_minlight", "").encode())

    def reflectivity(self):
        out = np.zeros((self.numtexinfo, 3), np.float32)
        self.L.oq_get_reflectivity(out.ctypes.data_as(C.c_void_p))
        return out

    def make_patches(self):
        self.L.oq_make_patches()
        self.L.oq_subdivide()
        return self.L.oq_num_patches()

    def create_direct_lights(self):  # lightmap.c:721-838
        self.L.oq_lights_from_patches()
        for e in self.entities:
            name = e.get("classname", "")
            if not name.startswith("light"):
                continue
            target = e.get("target", "")
            torigin = None
            if target:
                t = next((x for x in self.entities if x.get("targetname", "") == target), None)
                torigin = t.get("origin", "").encode() if t is not None else None
            b = lambda k: e.get(k, "").encode()
            self.L.oq_add_light_entity(name.encode(), b("origin"), b("light"), b("_light"), b("style"), b("_style"),
                                       b("_color"), target.encode(), torigin, b("_cone"), b("angle"))
        return self.L.oq_num_lights()

    def patches(self):
        n = self.L.oq_num_patches()
        a = dict(origin=np.zeros((n, 3), np.float32), normal=np.zeros((n, 3), np.float32), area=np.zeros(n, np.float32),
                 cluster=np.zeros(n, np.int32), totallight=np.zeros((n, 3), np.float32),
                 samplelight=np.zeros((n, 3), np.float32), samples=np.zeros(n, np.int32),
                 numtransfers=np.zeros(n, np.int32), face=np.zeros(n, np.int32), next=np.zeros(n, np.int32))
        self.L.oq_get_patches(*[a[k].ctypes.data_as(C.c_void_p) for k in
                                ("origin", "normal", "area", "cluster", "totallight", "samplelight", "samples",
                                 "numtransfers", "face", "next")])
        return a

    def build_facelights(self):
        self.L.oq_build_facelights()

    def facelight_style0(self):
        out = []
        for f in range(self.numfaces):
            ns = C.c_int()
            sty = (C.c_int * 32)()
            org = C.POINTER(C.c_float)()
            smp = C.POINTER(C.c_float)()
            nst = self.L.oq_facelight(f, C.byref(ns), sty, C.byref(org), C.byref(smp))
            if nst:
                out.append(np.ctypeslib.as_array(smp, shape=(ns.value, 3)).copy())
        return np.concatenate(out) if out else np.zeros((0, 3), np.float32)

    def make_transfers(self):
        self.L.oq_make_transfers()
        return self.L.oq_total_transfer()

    def transfers(self):
        nnz = self.L.oq_total_transfer()
        patch = np.zeros(nnz, np.uint32)
        tr = np.zeros(nnz, np.uint16)
        self.L.oq_get_transfers(patch.ctypes.data_as(C.c_void_p), tr.ctypes.data_as(C.c_void_p))
        return patch, tr

    def bounce(self):
        added = np.zeros(max(self.numbounce, 1), np.float32)
        self.L.oq_bounce(added.ctypes.data_as(C.c_void_p))
        return added[:self.numbounce]

    def final_light(self):
        cap = 0x200000
        out = np.zeros(cap, np.uint8)
        lightofs = np.zeros(self.numfaces, np.int32)
        styles = np.zeros((self.numfaces, 4), np.uint8)
        n = self.L.oq_final_light(out.ctypes.data_as(C.c_void_p), cap, lightofs.ctypes.data_as(C.c_void_p),
                                  styles.ctypes.data_as(C.c_void_p))
        if n < 0:
            raise RuntimeError("oracle: Error() in FinalLightFace")
        return out[:n].copy(), lightofs, styles

    def ray_counts(self):
        a, b = C.c_longlong(), C.c_longlong()
        self.L.oq_ray_counts(C.byref(a), C.byref(b))
        return a.value, b.value

    def test_lines(self, starts, stops):
        starts = np.ascontiguousarray(starts, np.float32)
        stops = np.ascontiguousarray(stops, np.float32)
        fp = C.POINTER(C.c_float)
        return np.array([self.L.oq_test_line(starts[i].ctypes.data_as(fp), stops[i].ctypes.data_as(fp))
                         for i in range(len(starts))], np.int32)

    def point_in_leaf(self, pts):
        """PointInLeafnum, qrad3.c:131-150, per point."""
        pts = np.ascontiguousarray(pts, np.float32)
        fp = C.POINTER(C.c_float)
        return np.array([self.L.oq_point_in_leaf(pts[i].ctypes.data_as(fp)) for i in range(len(pts))], np.int32)

    def rad_world(self):
        """RadWorld, qrad3.c:494-536. Returns (lump bytes, added)."""
        self.make_patches()
        self.create_direct_lights()
        self.build_facelights()
        added = np.zeros(0, np.float32)
        if self.numbounce > 0:
            self.make_transfers()
            added = self.bounce()
        lump, lightofs, styles = self.final_light()
        return lump, added
