"""Multi-GPU partition of a branch by continuation step.

PALC is a sequential recurrence (src/Continuation.jl:458-504), so one branch is cut into segments: rank r
advances `steps` continuation steps from a seed pair (z_a, z_b) = two consecutive points of a scout run, using the
reference's two-point start (iterate_from_two_points, src/Continuation.jl:408-456; same mechanism as branch
switching, src/bifdiagram/BranchSwitching.jl:8-44).  State vectors are replicated; the only collective is an
all_gather of the per-step rows (lambda, ||u||, itnewton, itlinear), 32 B per step.
"""
import numpy as np

ROW = ("param", "x", "itnewton", "itlinear")


def seed_steps(rank, stride):
    """Scout step indices whose states seed rank `rank` (first point, second point)."""
    return rank * stride, rank * stride + 1


class SeedGrabber:
    """Continuation callback that keeps the states at the scout steps this rank needs and stops the scout there."""

    def __init__(self, rank, stride, copy):
        self.a, self.b = seed_steps(rank, stride)
        self.copy, self.seeds = copy, {}
        self.last = []  # the two most recent scout points: fallback seeds if the branch ends before step b

    def __call__(self, st):
        if st.step in (self.a, self.b):
            self.seeds[st.step] = (self.copy(st.z_u), st.z_p)
        self.last = (self.last + [(st.step, self.copy(st.z_u), st.z_p)])[-2:]  # a device copy per scout step: negligible
        return st.step < self.b

    def pair(self):
        if self.a in self.seeds and self.b in self.seeds:
            (u0, p0), (u1, p1) = self.seeds[self.a], self.seeds[self.b]
            return u0, p0, u1, p1
        # the scout stopped early (parameter bound reached / ds < dsmin): seed from its last two points
        (_, u0, p0), (_, u1, p1) = self.last
        return u0, p0, u1, p1


def segment_starts(rank, world, steps, subs=4):
    """Interleaved partition of a window of world*steps scout steps into world*nsub sub-segments of `sub` steps:
    rank r owns sub-segments g = s*world + r (s = 0..nsub-1), i.e. every rank samples the whole window, so that hard
    stretches of the branch (folds of the snaking region: rejected steps) are spread over the ranks.
    Returns [(scout_step_of_first_seed, nsteps), ...]."""
    sub = max(1, steps // subs)
    out, left, s = [], steps, 0
    while left > 0:
        k = min(sub, left)
        out.append(((s * world + rank) * sub, k))
        left -= k
        s += 1
    return out


class MultiSeedGrabber:
    """Scout callback collecting the seed pairs (steps a, a+1) of several sub-segments; stops after the last one."""

    def __init__(self, starts, copy):
        self.want = sorted(set(a for a, _ in starts) | set(a + 1 for a, _ in starts))
        self.stop_at = self.want[-1]
        self.copy, self.seeds, self.last = copy, {}, []

    def __call__(self, st):
        if st.step in self.want:
            self.seeds[st.step] = (self.copy(st.z_u), st.z_p)
        self.last = (self.last + [(st.step, self.copy(st.z_u), st.z_p)])[-2:]
        return st.step < self.stop_at

    def pair(self, a):
        if a in self.seeds and a + 1 in self.seeds:
            (u0, p0), (u1, p1) = self.seeds[a], self.seeds[a + 1]
            return u0, p0, u1, p1
        (_, u0, p0), (_, u1, p1) = self.last  # the scout ended early: seed from its last two points
        return u0, p0, u1, p1


def rows_to_array(rows, nrows):
    """Fixed-size (nrows, 4) array for the collective; unused rows are NaN."""
    R = np.full((nrows, len(ROW)), np.nan)
    for i, r in enumerate(rows[:nrows]):
        R[i] = [r[k] for k in ROW]
    return R


def all_gather_rows(rows, nrows, dist, torch, device):
    """The path's only collective: every rank receives every rank's rows.  Returns (world, nrows, 4) ndarray."""
    t = torch.tensor(rows_to_array(rows, nrows), dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def merge_branch(gathered):
    """Concatenate the segments in rank order (the `_cat!` order of src/Results.jl:436-459), dropping padding and each
    segment's first row when it repeats the previous segment's last seed point."""
    out = []
    for seg in gathered:
        seg = seg[~np.isnan(seg[:, 0])]
        if out and len(seg) and abs(seg[0, 0] - out[-1][-1, 0]) < 1e-12:
            seg = seg[1:]
        if len(seg):
            out.append(seg)
    return np.concatenate(out) if out else np.zeros((0, len(ROW)))
