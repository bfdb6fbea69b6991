"""On-disk sample dumps — mirror of pipeline/utils/save.py:32-41 (`save_structures`: extended XYZ through ASE).

ASE and pymatgen are not in this image, so the extended-XYZ frames are written directly in the layout ASE's writer
produces for periodic structures: atom count, a comment line `Lattice="ax ay az bx by bz cx cy cz"
Properties=species:S:1:pos:R:3 pbc="T T T"`, then one `symbol x y z` line per atom (Cartesian, %16.8f).  The cell is
pymatgen's `Lattice.from_parameters` convention — the same one `lattice_params_to_matrix_torch` uses
(models/diffcsp/utils.py:68-96) — so files round-trip through `ase.io.read(..., format="extxyz")`."""
import math
import os

from ..rewards.elements import SYMBOLS


def _cell(lengths, angles):
    a, b, c = (float(v) for v in lengths)
    al, be, ga = (math.radians(float(v)) for v in angles)
    val = (math.cos(al) * math.cos(be) - math.cos(ga)) / (math.sin(al) * math.sin(be))
    gs = math.acos(max(-1.0, min(1.0, val)))
    return [[a * math.sin(be), 0.0, a * math.cos(be)],
            [-b * math.sin(al) * math.cos(gs), b * math.sin(al) * math.sin(gs), b * math.cos(al)],
            [0.0, 0.0, c]]


def _frame(item):
    """(symbols, cartesian positions, 3x3 cell) of a pymatgen Structure or a sampled crystal"""
    if hasattr(item, "lattice") and hasattr(item, "cart_coords"):          # pymatgen
        cell = [list(map(float, r)) for r in item.lattice.matrix]
        return [str(s) for s in item.species], [list(map(float, r)) for r in item.cart_coords], cell
    cell = _cell(item.lengths.reshape(-1).tolist(), item.angles.reshape(-1).tolist())
    pos = []
    for f in item.frac_coords.tolist():
        pos.append([sum(f[k] * cell[k][d] for k in range(3)) for d in range(3)])
    return [SYMBOLS[int(z)] for z in item.atom_types.reshape(-1).tolist()], pos, cell


def save_structures(structures, save_dir, filename):
    """pipeline/utils/save.py:32-41: write all structures to one extxyz file, return its path"""
    save_path = os.path.join(save_dir, filename)
    try:
        with open(save_path, "w") as fh:
            for item in structures:
                sym, pos, cell = _frame(item)
                fh.write("%d\n" % len(sym))
                fh.write('Lattice="%s" Properties=species:S:1:pos:R:3 pbc="T T T"\n'
                         % " ".join(repr(float(v)) for row in cell for v in row))
                for s, p in zip(sym, pos):
                    fh.write("%-2s %16.8f %16.8f %16.8f\n" % (s, p[0], p[1], p[2]))
        return save_path
    except IOError as e:
        print(f"Got error {e} writing the generated structures to disk.")


def read_extxyz(path):
    """minimal reader of the frames `save_structures` writes (tests, tooling): list of (symbols, positions, cell)"""
    import re
    frames = []
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = 0
    while i < len(lines):
        n = int(lines[i])
        lat = [float(v) for v in re.search(r'Lattice="([^"]*)"', lines[i + 1]).group(1).split()]
        cell = [lat[0:3], lat[3:6], lat[6:9]]
        sym, pos = [], []
        for ln in lines[i + 2:i + 2 + n]:
            t = ln.split()
            sym.append(t[0]), pos.append([float(v) for v in t[1:4]])
        frames.append((sym, pos, cell))
        i += 2 + n
    return frames
