"""Deterministic synthetic inputs for the qrad3 hot path.

Generates Quake II ``.map`` text (room grids with doorways, and a direct-light
stress hall) plus the two side inputs qrad3 needs next to a map -- the palette
``assets/pics/colormap.pcx`` and ``assets/textures/<name>.wal`` -- and compiles a
map to a vised ``.bsp`` with the *reference* ``qbsp3``/``qvis3`` binaries, so that
the reference qrad3 and libqrad_cuda light byte-identical BSPs (BASELINE.json
north_star).

Reference behaviour this respects:
  * qrad3 finds its assets relative to a path component named ``assets``
    (src/tools/common/cmdlib.c:187-217) and reads ``../pics/colormap.pcx`` and
    ``../textures/<texinfo.texture>.wal`` (src/tools/qrad3/patches.c:51-80).
  * every light carries an explicit ``_color`` (finding F4: lightmap.c:793-798).
  * format limits: MAX_MAP_ENTITIES 2048, MAX_MAP_FACES 65536, world +-4096
    (src/tools/common/qfiles.h:223-243).
"""
from __future__ import annotations

import os
import random
import shutil
import struct
import subprocess
from pathlib import Path

TEX_WALL = "b200/wall"
TEX_FLOOR = "b200/floor"
TEX_CEIL = "b200/ceil"
TEX_LAMP = "b200/lamp"


# ----------------------------------------------------------------------------
# side inputs: palette + textures
# ----------------------------------------------------------------------------
def synthetic_palette() -> bytes:
    """768-byte RGB palette: a smooth 6x6x6 colour cube, then a grey ramp."""
    pal = bytearray()
    for i in range(216):
        r, g, b = i // 36, (i // 6) % 6, i % 6
        pal += bytes((r * 51, g * 51, b * 51))
    for i in range(40):
        v = int(round(i * 255 / 39))
        pal += bytes((v, v, v))
    return bytes(pal)


def write_pcx(path: Path, width: int, height: int, pixels: bytes, palette: bytes) -> None:
    """8-bit RLE PCX v5, the only flavour LoadPCX accepts (lbmlib.c:341-400)."""
    hdr = struct.pack(
        "<BBBBHHHHHH48sBBHH58s",
        0x0A, 5, 1, 8, 0, 0, width - 1, height - 1, width, height,
        bytes(48), 0, 1, width, 1, bytes(58),
    )
    body = bytearray()
    for y in range(height):
        row = pixels[y * width:(y + 1) * width]
        x = 0
        while x < width:
            v = row[x]
            run = 1
            while x + run < width and row[x + run] == v and run < 63:
                run += 1
            if run > 1 or (v & 0xC0) == 0xC0:
                body += bytes((0xC0 | run, v))
            else:
                body.append(v)
            x += run
    path.write_bytes(hdr + bytes(body) + b"\x0c" + palette)


def write_wal(path: Path, name: str, texels: bytes, width: int, height: int,
              flags: int = 0, contents: int = 0, value: int = 0) -> None:
    """miptex_t header (qfiles.h:196-205) + 4 mip levels (only mip 0 is read by qrad3)."""
    mips = [texels]
    w, h = width, height
    for _ in range(3):
        w, h = max(1, w // 2), max(1, h // 2)
        mips.append(bytes(w * h))
    ofs = []
    pos = 100
    for m in mips:
        ofs.append(pos)
        pos += len(m)
    hdr = struct.pack("<32sII4I32siii", name.encode(), width, height, *ofs, b"", flags, contents, value)
    assert len(hdr) == 100
    path.write_bytes(hdr + b"".join(mips))


def _texels(seed: int, width: int, height: int, lo: int, hi: int) -> bytes:
    rng = random.Random(seed)
    return bytes(rng.randint(lo, hi) for _ in range(width * height))


def write_assets(root: Path) -> Path:
    """Create ``root/assets/{pics,textures/b200,textures/eq2,maps}`` with synthetic content.

    ``eq2/*`` are stand-ins with the names assets/maps/sample.map uses, so the
    sample map can be lit without the reference's art files.
    """
    assets = Path(root) / "assets"
    (assets / "pics").mkdir(parents=True, exist_ok=True)
    (assets / "maps").mkdir(parents=True, exist_ok=True)
    pal = synthetic_palette()
    write_pcx(assets / "pics" / "colormap.pcx", 16, 16, bytes(range(256)), pal)
    specs = {
        TEX_WALL: (11, 120, 215), TEX_FLOOR: (12, 40, 150), TEX_CEIL: (13, 216, 255), TEX_LAMP: (14, 180, 255),
        "eq2/wall": (21, 100, 215), "eq2/floor": (22, 30, 160), "eq2/ceiling": (23, 216, 250),
    }
    for name, (seed, lo, hi) in specs.items():
        p = assets / "textures" / (name + ".wal")
        p.parent.mkdir(parents=True, exist_ok=True)
        write_wal(p, name, _texels(seed, 16, 16, lo, hi), 16, 16)
    return assets


def link_assets(dst_root: Path, src_assets: Path) -> Path:
    """Make ``dst_root/assets`` with pics/ and textures/ copied from ``src_assets``; returns maps dir."""
    dst = Path(dst_root) / "assets"
    (dst / "maps").mkdir(parents=True, exist_ok=True)
    for sub in ("pics", "textures"):
        if not (dst / sub).exists():
            shutil.copytree(Path(src_assets) / sub, dst / sub)
    return dst / "maps"


# ----------------------------------------------------------------------------
# .map text
# ----------------------------------------------------------------------------
def _face(p0, p1, p2, tex, flags=0, value=0, contents=0) -> str:
    def pt(p):
        return "( %d %d %d )" % p
    return "%s %s %s %s 0 0 0 1 1 %d %d %d\n" % (pt(p0), pt(p1), pt(p2), tex, contents, flags, value)


def box_brush(x0, y0, z0, x1, y1, z1, tex, top_tex=None, bottom_tex=None, top_light=0, bottom_light=0,
              contents=0) -> str:
    """Axis-aligned brush; plane point order as written by Quake II editors (cf. assets/maps/sample.map)."""
    top_tex = top_tex or tex
    bottom_tex = bottom_tex or tex
    s = "{\n"
    c = contents
    s += _face((x0, y0, z0), (x0, y0 + 1, z0), (x0, y0, z0 + 1), tex, contents=c)
    s += _face((x1, y1, z1), (x1, y1, z1 + 1), (x1, y1 + 1, z1), tex, contents=c)
    s += _face((x0, y0, z0), (x0, y0, z0 + 1), (x0 + 1, y0, z0), tex, contents=c)
    s += _face((x1, y1, z1), (x1 + 1, y1, z1), (x1, y1, z1 + 1), tex, contents=c)
    s += _face((x1, y1, z1), (x1, y1 + 1, z1), (x1 + 1, y1, z1), top_tex, 1 if top_light else 0, top_light, contents=c)
    s += _face((x0, y0, z0), (x0 + 1, y0, z0), (x0, y0 + 1, z0), bottom_tex, 1 if bottom_light else 0, bottom_light,
               contents=c)
    s += "}\n"
    return s


def _plane(a, u, v, tex, flags=0, value=0) -> str:
    """Face through point a whose outward normal is cross(u, v) (Quake II order: normal = cross(p0 - p1, p2 - p1))."""
    p0 = (a[0] + u[0], a[1] + u[1], a[2] + u[2])
    p2 = (a[0] + v[0], a[1] + v[1], a[2] + v[2])
    return _face(p0, a, p2, tex, flags, value)


def diamond_pillar(cx, cy, r, z0, z1, tex) -> str:
    """Vertical pillar whose four walls are 45-degree (non-axial, PLANE_ANYX/ANYY) planes."""
    s = "{\n"
    s += _plane((cx + r, cy, z0), (0, 0, 1), (1, -1, 0), tex)     # normal ( 1,  1, 0)
    s += _plane((cx - r, cy, z0), (0, 0, 1), (-1, 1, 0), tex)     # normal (-1, -1, 0)
    s += _plane((cx + r, cy, z0), (0, 0, 1), (-1, -1, 0), tex)    # normal ( 1, -1, 0)
    s += _plane((cx - r, cy, z0), (0, 0, 1), (1, 1, 0), tex)      # normal (-1,  1, 0)
    s += _face((cx, cy, z1), (cx, cy + 1, z1), (cx + 1, cy, z1), tex)
    s += _face((cx, cy, z0), (cx + 1, cy, z0), (cx, cy + 1, z0), tex)
    return s + "}\n"


def ramp_brush(x0, y0, z0, x1, y1, h0, h1, tex, top_flags=0) -> str:
    """Box footprint [x0,x1]x[y0,y1] from z0 up to a sloped top (z0+h0 at x0, z0+h1 at x1): a PLANE_ANYZ plane."""
    s = "{\n"
    s += _face((x0, y0, z0), (x0, y0 + 1, z0), (x0, y0, z0 + 1), tex)
    s += _face((x1, y1, z0), (x1, y1, z0 + 1), (x1, y1 + 1, z0), tex)
    s += _face((x0, y0, z0), (x0, y0, z0 + 1), (x0 + 1, y0, z0), tex)
    s += _face((x1, y1, z0), (x1 + 1, y1, z0), (x1, y1, z0 + 1), tex)
    s += _plane((x0, y0, z0 + h0), (x1 - x0, 0, h1 - h0), (0, 1, 0), tex, top_flags)
    s += _face((x0, y0, z0), (x0 + 1, y0, z0), (x0, y0 + 1, z0), tex)
    return s + "}\n"


def _light_entity(rng, x, y, z, spot=False) -> str:
    col = tuple(round(rng.uniform(0.5, 1.0), 3) for _ in range(3))
    s = '{\n"classname" "%s"\n' % ("light_spot" if spot else "light")
    s += '"origin" "%d %d %d"\n' % (x, y, z)
    s += '"light" "%d"\n' % rng.randint(200, 400)
    s += '"_color" "%g %g %g"\n' % col
    if spot:
        s += '"_cone" "%d"\n' % rng.randint(10, 40)
        s += '"angle" "%d"\n' % rng.choice([-2, -2, -1, 0, 45, 90, 135, 180, 225, 270, 315])
    s += "}\n"
    return s


def room_grid_map(nx: int, ny: int, seed: int = 1, room: int = 256, height: int = 128, wall: int = 16,
                  door_w: int = 64, door_h: int = 96, clutter: float = 0.0, lamps: float = 0.0) -> str:
    """nx x ny rooms joined by doorways, one point light per room.

    clutter: probability of a free-standing box in a room (extra occluders).
    lamps:   probability that a room gets an emissive (SURF_LIGHT) ceiling panel.
    """
    rng = random.Random(seed)
    P = room + wall
    X, Y = nx * P + wall, ny * P + wall
    ox = -((X // 2) // 16) * 16
    oy = -((Y // 2) // 16) * 16
    oz = -((height // 2) // 16) * 16
    if max(X, Y) > 8000:
        raise ValueError("map exceeds the +-4096 world bound")
    out = ["// generated by qengine_b200.mapgen.room_grid_map(%d, %d, seed=%d)\n" % (nx, ny, seed),
           '{\n"classname" "worldspawn"\n']

    def B(x0, y0, z0, x1, y1, z1, tex, **kw):
        out.append(box_brush(x0 + ox, y0 + oy, z0 + oz, x1 + ox, y1 + oy, z1 + oz, tex, **kw))

    B(0, 0, -wall, X, Y, 0, TEX_WALL, top_tex=TEX_FLOOR)
    B(0, 0, height, X, Y, height + wall, TEX_WALL, bottom_tex=TEX_CEIL)
    # outer walls
    B(0, 0, 0, wall, Y, height, TEX_WALL)
    B(X - wall, 0, 0, X, Y, height, TEX_WALL)
    B(0, 0, 0, X, wall, height, TEX_WALL)
    B(0, Y - wall, 0, X, Y, height, TEX_WALL)
    # pillars at interior crossings
    for k in range(1, nx):
        for l in range(1, ny):
            B(k * P, l * P, 0, k * P + wall, l * P + wall, height, TEX_WALL)
    d0 = wall + (room - door_w) // 2
    d1 = d0 + door_w
    # interior walls perpendicular to x (between room i-1 and i), one segment per row j
    for k in range(1, nx):
        for j in range(ny):
            y0 = j * P
            B(k * P, y0 + wall, 0, k * P + wall, y0 + d0, height, TEX_WALL)
            B(k * P, y0 + d1, 0, k * P + wall, y0 + P, height, TEX_WALL)
            B(k * P, y0 + d0, door_h, k * P + wall, y0 + d1, height, TEX_WALL)
    for l in range(1, ny):
        for i in range(nx):
            x0 = i * P
            B(x0 + wall, l * P, 0, x0 + d0, l * P + wall, height, TEX_WALL)
            B(x0 + d1, l * P, 0, x0 + P, l * P + wall, height, TEX_WALL)
            B(x0 + d0, l * P, door_h, x0 + d1, l * P + wall, height, TEX_WALL)
    ents = []
    for j in range(ny):
        for i in range(nx):
            cx, cy = i * P + wall + room // 2, j * P + wall + room // 2
            if clutter and rng.random() < clutter:
                bx = i * P + wall + 16 * rng.randint(1, room // 16 - 4)
                by = j * P + wall + 16 * rng.randint(1, room // 16 - 4)
                bw, bd, bh = 16 * rng.randint(1, 3), 16 * rng.randint(1, 3), 16 * rng.randint(1, 4)
                if not (bx <= cx <= bx + bw and by <= cy <= by + bd):
                    B(bx, by, 0, bx + bw, by + bd, bh, TEX_WALL)
            if lamps and rng.random() < lamps:
                lx = i * P + wall + 16 * rng.randint(1, room // 16 - 3)
                ly = j * P + wall + 16 * rng.randint(1, room // 16 - 3)
                B(lx, ly, height - 8, lx + 32, ly + 32, height, TEX_WALL, bottom_tex=TEX_LAMP,
                  bottom_light=rng.randint(300, 900))
            jx, jy = rng.randint(-48, 48), rng.randint(-48, 48)
            ents.append(_light_entity(rng, cx + jx + ox, cy + jy + oy, height - 32 + oz))
    out.append("}\n")
    out.append('{\n"classname" "info_player_start"\n"origin" "%d %d %d"\n}\n'
               % (wall + room // 2 + ox + 32, wall + room // 2 + oy + 32, 24 + oz))
    out.extend(ents)
    return "".join(out)


def light_hall_map(nlights: int = 2000, seed: int = 2, size: int = 2048, height: int = 256, wall: int = 16,
                   pillars: int = 7) -> str:
    """One open hall with a pillar lattice and many light / light_spot entities."""
    if nlights > 2040:
        raise ValueError("MAX_MAP_ENTITIES is 2048 (qfiles.h:225)")
    rng = random.Random(seed)
    X = Y = size + 2 * wall
    ox = oy = -((X // 2) // 16) * 16
    oz = -((height // 2) // 16) * 16
    out = ["// generated by qengine_b200.mapgen.light_hall_map(%d, seed=%d)\n" % (nlights, seed),
           '{\n"classname" "worldspawn"\n']

    def B(x0, y0, z0, x1, y1, z1, tex, **kw):
        out.append(box_brush(x0 + ox, y0 + oy, z0 + oz, x1 + ox, y1 + oy, z1 + oz, tex, **kw))

    B(0, 0, -wall, X, Y, 0, TEX_WALL, top_tex=TEX_FLOOR)
    B(0, 0, height, X, Y, height + wall, TEX_WALL, bottom_tex=TEX_CEIL)
    B(0, 0, 0, wall, Y, height, TEX_WALL)
    B(X - wall, 0, 0, X, Y, height, TEX_WALL)
    B(0, 0, 0, X, wall, height, TEX_WALL)
    B(0, Y - wall, 0, X, Y, height, TEX_WALL)
    pitch = size // (pillars + 1)
    boxes = []
    for a in range(1, pillars + 1):
        for b in range(1, pillars + 1):
            px, py = wall + a * pitch - 32, wall + b * pitch - 32
            B(px, py, 0, px + 64, py + 64, height, TEX_WALL)
            boxes.append((px, py, px + 64, py + 64))
    out.append("}\n")
    out.append('{\n"classname" "info_player_start"\n"origin" "%d %d %d"\n}\n' % (wall + 64 + ox, wall + 64 + oy, 24 + oz))
    n = 0
    while n < nlights:
        x = rng.randint(wall + 24, X - wall - 24)
        y = rng.randint(wall + 24, Y - wall - 24)
        z = rng.randint(24, height - 24)
        if any(b[0] - 8 <= x <= b[2] + 8 and b[1] - 8 <= y <= b[3] + 8 for b in boxes):
            continue
        out.append(_light_entity(rng, x + ox, y + oy, z + oz, spot=(n % 2 == 1)))
        n += 1
    return "".join(out)


# ----------------------------------------------------------------------------
# compile with the reference map tools
# ----------------------------------------------------------------------------
def find_map_tools(repo_root: Path | None = None) -> Path | None:
    """Directory holding the reference-built ``qbsp3`` and ``qvis3`` (oracle/Makefile target ``ref``)."""
    root = Path(repo_root) if repo_root else Path(__file__).resolve().parent.parent
    d = root / "oracle" / "_ref"
    if (d / "qbsp3").exists() and (d / "qvis3").exists():
        return d
    return None


def compile_map(map_path: Path, tools: Path, chop: int | None = None, fastvis: bool = False,
                quiet: bool = True) -> Path:
    """Run reference qbsp3 then qvis3 on ``map_path`` (must live under .../assets/maps/)."""
    map_path = Path(map_path).resolve()
    cmd = [str(tools / "qbsp3")]
    if chop:
        cmd += ["-chop", str(chop)]
    cmd.append(str(map_path))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 or "LEAKED" in r.stdout:
        raise RuntimeError("qbsp3 failed:\n" + r.stdout[-2000:] + r.stderr[-500:])
    bsp = map_path.with_suffix(".bsp")
    cmd = [str(tools / "qvis3")] + (["-fast"] if fastvis else []) + [str(bsp)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("qvis3 failed:\n" + r.stdout[-2000:] + r.stderr[-500:])
    if not quiet:
        print(r.stdout[-400:])
    return bsp


def build_case(workdir: Path, name: str, map_text: str, tools: Path, chop: int | None = None,
               fastvis: bool = False) -> Path:
    """Write synthetic assets + ``name.map`` under ``workdir/assets`` and compile it; returns the .bsp path."""
    assets = write_assets(workdir)
    mp = assets / "maps" / (name + ".map")
    mp.write_text(map_text)
    return compile_map(mp, tools, chop=chop, fastvis=fastvis)
