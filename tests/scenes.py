"""Seeded scene generators shared by the CPU and GPU parity tests."""
import numpy as np


def random_scene(pm, seed, width, height, n_items, mode, rules=False):
    """Mixed Fill / Poly / Line / Circle items.  `mode` picks the coordinate lattice: 'int16'
    puts vertices on multiples of 8 (tile corners and edges: knife-edge cull decisions), 'int' on
    integers, 'half' on half-integers, 'float' anywhere."""
    rng = np.random.default_rng(seed)
    enc = pm.Encoder(1 << 20)
    enc.begin_group(n_items)

    def pt():
        if mode == "int16":
            return (float(rng.integers(-2, width // 8 + 3) * 8), float(rng.integers(-2, height // 8 + 3) * 8))
        if mode == "int":
            return (float(rng.integers(-20, width + 20)), float(rng.integers(-20, height + 20)))
        if mode == "half":
            return (rng.integers(-20, 2 * width + 40) / 2.0, rng.integers(-20, 2 * height + 40) / 2.0)
        return (rng.uniform(-30, width + 30), rng.uniform(-30, height + 30))

    for _ in range(n_items):
        kind = rng.integers(0, 10)
        alpha = 255 if rng.random() < 0.6 else int(rng.integers(1, 255))
        rgba = (int(rng.integers(0, 1 << 24)) << 8) | alpha
        if kind < 6:
            n = int(rng.integers(1, 9))
            pts = [pt() for _ in range(n)]
            if rng.random() < 0.3 and n >= 2:
                pts[1] = (pts[1][0], pts[0][1])  # horizontal edge
            if rng.random() < 0.3 and n >= 3:
                pts[2] = (pts[1][0], pts[2][1])  # vertical edge
            if rules:  # extension: the item's flags word picks the fill rule (PM_FLAG_FILL_RULES), sometimes with a second subpath
                fl = int(rng.integers(0, 2))
                if rng.random() < 0.4:
                    enc.fill_subpaths([pts, [pt() for _ in range(int(rng.integers(3, 6)))]], rgba, flags=fl)
                else:
                    enc.fill(np.array(pts), rgba, flags=fl)
            else:
                enc.fill(np.array(pts), rgba)
        elif kind < 8:
            n = int(rng.integers(1, 9))
            enc.polyline(np.array([pt() for _ in range(n)]), rgba, float(rng.choice([0.7, 1.0, 2.0, 5.5, 17.0])))
        elif kind == 8:
            enc.stroke_line(pt(), pt(), float(rng.choice([0.7, 2.0, 9.0])), rgba)
        else:
            c = pt()
            enc.circle(c[0], c[1], float(rng.uniform(1, 40)))
    enc.end_group()
    return enc.bytes()


FUZZ_MODES = ["int16", "int", "half", "float"]


def fuzz_case(pm, seed):
    rng = np.random.default_rng(1000 + seed)
    mode = FUZZ_MODES[seed % 4]
    width = int(rng.choice([16, 48, 100, 256, 300, 520]))
    height = int(rng.choice([16, 40, 64, 130, 272]))
    scene = random_scene(pm, seed, width, height, int(rng.integers(1, 25)), mode)
    return scene, width, height, int(seed % 7 == 0)


def rules_case(pm, seed):
    """Fuzz scene whose Fill items carry random fill-rule flags and second subpaths (for PM_FLAG_FILL_RULES)."""
    rng = np.random.default_rng(5000 + seed)
    mode = FUZZ_MODES[1 + seed % 3]
    width = int(rng.choice([48, 100, 256, 300]))
    height = int(rng.choice([40, 64, 130, 272]))
    return random_scene(pm, 7000 + seed, width, height, int(rng.integers(1, 20)), mode, rules=True), width, height


def items_equal(a, b):
    return (np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["items"], b["items"])
            and np.array_equal(a["solid"], b["solid"]))


def stacked_scene(pm, n_layers, width=96, height=64, seed=7):
    """Many translucent layers over the same tiles: hundreds of records per tile (overflow chain,
    shared-memory index cap of the fill kernel), with an opaque cover in the middle of the stack."""
    rng = np.random.default_rng(seed)
    enc = pm.Encoder(1 << 22)
    enc.begin_group(n_layers)
    for i in range(n_layers):
        if i == n_layers // 3:
            enc.fill(np.array([(-8.0, -8.0), (width + 8.0, -8.0), (width + 8.0, height / 2.0), (-8.0, height / 2.0)]), 0x204080ff)
            continue
        x0, y0 = rng.uniform(-20, width / 2), rng.uniform(-20, height / 2)
        x1, y1 = x0 + rng.uniform(10, width), y0 + rng.uniform(10, height)
        rgba = (int(rng.integers(0, 1 << 24)) << 8) | int(rng.integers(8, 60))
        if i % 5 == 0:
            enc.polyline(np.array([(x0, y0), (x1, y1), (x0, y1)]), rgba, 3.0)
        else:
            enc.fill(np.array([(x0, y0), (x1, y0 + 3.0), (x1, y1), (x0 - 2.0, y1)]), rgba)
    enc.end_group()
    return enc.bytes()
