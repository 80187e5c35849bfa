"""`addplot!` / `plot!` mirrors (src/lib.rs:1163-1240, `PlotData` :490-600): named plots holding named series
of (x, y) points, filled from device-side reductions so that recording a series does not download the
population.  Rendering (the reference draws PNGs with `plotters` and TUI tabs) stays with the reference;
`PlotData.to_csv` writes the points."""
from collections import OrderedDict

DATA = OrderedDict()        # the reference's global `DATA: Mutex<HashMap<String, PlotData>>` (lib.rs:680)


class PlotData:
    """lib.rs:491-528"""

    def __init__(self, name, xlabel, ylabel, to_be_stored=False):
        self.name, self.xlabel, self.ylabel, self.to_be_stored = name, xlabel, ylabel, to_be_stored
        self.series = OrderedDict()
        self.min_x = self.min_y = float("inf")
        self.max_x = self.max_y = float("-inf")

    def add_point(self, series, x, y):
        self.series.setdefault(series, []).append((float(x), float(y)))
        self.min_x, self.max_x = min(self.min_x, x), max(self.max_x, x)
        self.min_y, self.max_y = min(self.min_y, y), max(self.max_y, y)

    def to_csv(self, path):
        with open(path, "w") as f:
            f.write(f"series,{self.xlabel},{self.ylabel}\n")
            for name, pts in self.series.items():
                for x, y in pts:
                    f.write(f"{name},{x!r},{y!r}\n")


def addplot(name, xlabel, ylabel, to_be_stored=False):
    """addplot!(name, xlabel, ylabel [, to_be_stored])  lib.rs:1163-1200"""
    DATA[name] = PlotData(name, xlabel, ylabel, to_be_stored)
    return DATA[name]


def plot(name, series, x, y):
    """plot!(name, series, x, y)  lib.rs:1202-1240 — the plot must exist (the reference panics otherwise)"""
    if name not in DATA:
        raise KeyError(f"plot {name!r} does not exist: use addplot first")
    DATA[name].add_point(series, x, y)


def plot_series(name, series, field, params, nsteps, every, y_of, x0=0):
    """Run `nsteps` device steps of `field` and add one point per `every` steps to plot `name`: x = step
    number, y = y_of(row) with row the dict of device-side sums after that step (Field2D.run_boids_series).
    What a model's `after_step` does with `plot!`, without a host round trip per step."""
    keys = ("sum_x", "sum_y", "sum_ldx", "sum_ldy", "sum_speed", "sum_xx", "sum_yy")
    rows = field.run_boids_series(params, nsteps, every)
    n = field.num_objects()
    for r, row in enumerate(rows):
        red = {k: float(row[i]) for i, k in enumerate(keys)}
        red["n"] = n
        plot(name, series, x0 + (r + 1) * every, y_of(red))
    return rows
