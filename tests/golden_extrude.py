"""The reference's golden 10 x 1 x 10 (Periodic, Flat, Bounded) case extruded into 3-D, so that the
reference's own pinned vectors (test/data/reference_sim.jld2 -> reference_offline_output.jld2) also pin
the code no reference test exercises: y-direction advection, 3-D grids, the production TMA kernel.

  mode "y-replicated":  x = golden x, the golden fields replicated along a Periodic y axis of `n_rep`
                        cells, v absent (zero): every y-column of every output must equal the golden field.
  mode "x-to-y":        the permuted twin: golden x mapped onto y (u -> v, xi_u -> xi_v), x replicated:
                        every x-column must equal the golden field.

Pure numpy: used by the CPU oracle tests and by the GPU parity tests alike.
"""
import numpy as np

H = 3
NX, NZ = 10, 10


def rehalo_x(a, h_new):
    """Periodic x axis of a golden parent array (16, 1, nz) re-haloed to width h_new."""
    core = a[H:H + NX]
    return np.concatenate([core[NX - h_new:], core, core[:h_new]], axis=0)


def rehalo_z(a, h_new, face):
    """Bounded z axis re-haloed: the first halo cell mirrors the interior (zero-flux default BC), the outer
    ones are zero; face fields keep zeros beyond the wall levels."""
    n = NZ + 1 if face else NZ
    core = a[:, :, H:H + n]
    lo = np.zeros(a.shape[:2] + (h_new,), dtype=a.dtype)
    hi = np.zeros(a.shape[:2] + (h_new,), dtype=a.dtype)
    if not face:
        lo[:, :, -1] = core[:, :, 0]
        hi[:, :, 0] = core[:, :, -1]
    return np.concatenate([lo, core, hi], axis=2)


def extrude(golden_sim, mode, n_rep=8, halo=3):
    """Returns (grid_kwargs, times, fields, vel_names, rename) for the extruded case; `rename` maps the extruded
    run's output names to the golden file's names."""
    nt = len(golden_sim["t"])
    times = [float(t) for t in golden_sim["t"]]
    h = halo
    fields = {}

    def prep(name, n):
        a = np.asarray(golden_sim[name][n])
        a = rehalo_z(rehalo_x(a, h), h, face=(name == "w"))
        return a                                            # (10 + 2h, 1, nz + 2h)

    if mode == "y-replicated":
        vel_names, rename = ("u", "w"), {}
        for name in ("u", "w", "b"):
            fields[name] = [np.asfortranarray(np.repeat(prep(name, n), n_rep + 2 * h, axis=1)) for n in range(nt)]
        grid = dict(size=(NX, n_rep, NZ), x=(-5000, 5000), y=(0, 1000.0 * n_rep), z=(-100, 0),
                    topology=("Periodic", "Periodic", "Bounded"), halo=(h, h, h))
    elif mode == "x-to-y":
        vel_names = ("v", "w")
        rename = {"xi_v": "xi_u", "v_Lagrangian_filtered": "u_Lagrangian_filtered"}
        src = {"v": "u", "w": "w", "b": "b"}
        for name in ("v", "w", "b"):
            fields[name] = [np.asfortranarray(np.repeat(np.transpose(prep(src[name], n), (1, 0, 2)), n_rep + 2 * h, axis=0))
                            for n in range(nt)]
        grid = dict(size=(n_rep, NX, NZ), x=(0, 1000.0 * n_rep), y=(-5000, 5000), z=(-100, 0),
                    topology=("Periodic", "Periodic", "Bounded"), halo=(h, h, h))
    else:
        raise ValueError(mode)
    return grid, times, fields, vel_names, rename


def columns_of(arr, mode, halo=3):
    """Every replicated column of a 3-D output parent array as golden-shaped (10 + 2h, 1, nz + 2h) arrays (interior
    replicas only)."""
    h = halo
    if mode == "y-replicated":
        return [arr[:, j:j + 1, :] for j in range(h, arr.shape[1] - h)]
    return [np.transpose(arr[i:i + 1, :, :], (1, 0, 2)) for i in range(h, arr.shape[0] - h)]


def golden_rehaloed(a, name, halo=3):
    """A golden OUTPUT array (centre field with the App. A.6 halo pattern) at another halo width."""
    if halo == H:
        return np.asarray(a)
    x = rehalo_x(np.asarray(a), halo)
    core = x[:, :, H:H + NZ]
    lo = np.zeros(x.shape[:2] + (halo,)); hi = np.zeros(x.shape[:2] + (halo,))
    lo[:, :, -1] = core[:, :, 0]; hi[:, :, 0] = core[:, :, -1]
    return np.concatenate([lo, core, hi], axis=2)
