"""Checkpoint / resume of a progressive render (SURVEY.md aux subsystems; the reference keeps this state only in memory:
its accumulation buffer and `accumulations` count, Renderer.cpp:200-205,1262).

The state is the selected accumulation target's double4 sums (bpt_read_accumulation) plus the index of the next sample. A sample
is a pure function of (pixel, accumulation index, scene), so `render(first_sample=next_sample)` after `load` continues bit for
bit as if the render had never stopped - on the same GPU, on another one, or split over several (each resumes its own range)."""
import numpy as np

FORMAT = 1


def save(ctx, path, next_sample, **meta):
    """Writes the selected target of `ctx` (a capi.Bpt) and the index of the next sample to render to `path` (.npz)."""
    sums = ctx.read_accumulation()
    np.savez(path, format=np.int32(FORMAT), sums=sums, next_sample=np.uint32(next_sample),
             **{f"meta_{k}": np.asarray(v) for k, v in meta.items()})
    return sums.shape[1], sums.shape[0]


def load(ctx, path):
    """Restores the selected target of `ctx` from `path`; returns (next_sample, meta)."""
    with np.load(path) as f:
        if int(f["format"]) != FORMAT:
            raise ValueError(f"{path}: checkpoint format {int(f['format'])}, expected {FORMAT}")
        sums = f["sums"]
        if sums.ndim != 3 or sums.shape[2] != 4 or sums.dtype != np.float64:
            raise ValueError(f"{path}: not an accumulation state")
        ctx.write_accumulation(sums)
        meta = {k[5:]: f[k] for k in f.files if k.startswith("meta_")}
        return int(f["next_sample"]), meta
