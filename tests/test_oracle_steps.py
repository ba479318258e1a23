"""The C oracle's STEP ORCHESTRATION against a second, independent restatement written in plain Python loops straight from
the reference source (small systems only).  The oracle header says "parity unpinned" because Julia cannot run here;
these tests make sure that at least two independent readings of the reference text agree on whole steps:
cell binning, half-stencil pair enumeration with Newton-3 scatter, the two force passes of update_verlet! (the second
on STALE cells), update_szabo! (with its sqrt(|vx|+|vy|) "speed"), update_rtp!, periodic / rigid walls! and update_time!.
"""
import math
from fractions import Fraction

import numpy as np
import pytest

import helpers as H

pkg = H.pkg


class RefSim:
    """Plain-Python restatement.  Every method cites the reference function it follows (1-based ids inside)."""

    def __init__(self, case, noise=None):
        st = case["mk"]()
        self.pos = [np.array(p, dtype=np.float64) for p in st.pos]
        self.second = ([np.array(v, dtype=np.float64) for v in st.vel] if hasattr(st, "vel") else [float(a) for a in st.pol_angle])
        self.n = len(self.pos)
        self.forces = [np.zeros(2) for _ in range(self.n)]
        g = case["geom"]
        self.size = np.array([g.length, g.height])
        self.bl = np.array([0.0, 0.0])
        self.periodic = isinstance(case["space"].wall_type, pkg.PeriodicWalls)
        cc = case["int_cfg"].chunks_cfg
        self.num_cols, self.num_rows = cc.num_cols, cc.num_rows
        self.chunk_l, self.chunk_h = g.length / cc.num_cols, g.height / cc.num_rows   # src/chunks.jl:27-28
        self.dyn = case["dyn"]
        self.dt = case["int_cfg"].dt
        self.time, self.num_steps = 0.0, 0
        self.neighbors = self.get_neighbors()
        self.chunks = None
        self.noise = noise
        self.radius = pkg.particle_radius(self.dyn)

    # src/chunks.jl:61-87 (periodic) — rows/cols 1-based, half stencil in the reference's order
    def get_neighbors(self):
        assert self.periodic, "the walled table is checked cell by cell in test_oracle_kat.py"
        R, C = self.num_rows, self.num_cols

        def get_id(x, n):
            if x == 0:
                return n
            if x % (n + 1) == 0:
                return 1
            return x

        return {(i, j): [(get_id(i + 1, R), get_id(j, C)), (get_id(i + 1, R), get_id(j + 1, C)),
                         (get_id(i, R), get_id(j + 1, C)), (get_id(i - 1, R), get_id(j + 1, C))]
                for i in range(1, R + 1) for j in range(1, C + 1)}

    # src/chunks.jl:120-163; Base.div(x, y) = trunc of the exact quotient
    def update_chunks(self):
        self.chunks = {(i, j): [] for i in range(1, self.num_rows + 1) for j in range(1, self.num_cols + 1)}
        for i in range(self.n):
            x, y = self.pos[i]
            ty = -y + self.bl[1] + self.size[1]
            tx = x - self.bl[0]
            row = int(Fraction(float(ty)) / Fraction(float(self.chunk_h))) + 1
            col = int(Fraction(float(tx)) / Fraction(float(self.chunk_l))) + 1
            row -= 1 if row == self.num_rows + 1 else 0
            col -= 1 if col == self.num_cols + 1 else 0
            self.chunks[(row, col)].append(i)

    # src/integration.jl:38-48
    def calc_diff(self, r1, r2):
        dr = r1 - r2
        if self.periodic:
            dr = dr - (np.abs(dr) > (self.size / 2)) * np.copysign(self.size, dr)
        return dr

    # src/integration.jl:62-109, src/configs.jl:354-397
    def calc_interaction(self, i, j):
        d = self.dyn
        dr = self.calc_diff(self.pos[i], self.pos[j])
        dist = math.sqrt(dr[0] ** 2 + dr[1] ** 2)
        if isinstance(d, pkg.LenJonesCfg):
            fmod = 4 * d.epsilon * (12 * d.sigma ** 12 / dist ** 13 - 6 * d.sigma ** 6 / dist ** 7)
            return fmod / dist * dr
        if isinstance(d, pkg.HarmTruncCfg):
            if dist > d.dist_max:
                return np.zeros(2)
            k = d.k_rep if dist < d.dist_eq else d.k_atr
            return (-k * (dist / d.dist_eq - 1)) / dist * dr
        if isinstance(d, pkg.SzaboCfg):
            if dist > d.r_max:
                return np.zeros(2)
            f_mod = d.k_adh / d.r_eq if dist > d.r_eq else d.k_rep / (d.r_max - d.r_eq)
            return -f_mod * (dist - d.r_eq) * dr
        cutoff = 2 ** (1 / 6) * d.sigma
        if dist > cutoff:
            return np.zeros(2)
        fmod = -4 * d.epsilon * (-12 * d.sigma ** 12 / dist ** 13 + 6 * d.sigma ** 6 / dist ** 7)
        return fmod / dist * dr

    # src/integration.jl:112-157 (col outer, row inner; same chunk j > i; then every particle of each stencil chunk)
    def calc_forces(self):
        for col in range(1, self.num_cols + 1):
            for row in range(1, self.num_rows + 1):
                chunk = self.chunks[(row, col)]
                for a, p1 in enumerate(chunk):
                    for p2 in chunk[a + 1:]:
                        f = self.calc_interaction(p1, p2)
                        self.forces[p1] = self.forces[p1] + f
                        self.forces[p2] = self.forces[p2] - f
                    for nb in self.neighbors[(row, col)]:
                        for p2 in self.chunks[nb]:
                            f = self.calc_interaction(p1, p2)
                            self.forces[p1] = self.forces[p1] + f
                            self.forces[p2] = self.forces[p2] - f

    def clean_forces(self):
        self.forces = [np.zeros(2) for _ in range(self.n)]

    # src/integration.jl:415-431 — pass 2 reuses the chunks of the un-drifted positions
    def update_verlet(self):
        old = [f.copy() for f in self.forces]
        term = self.dt ** 2 / 2
        for i in range(self.n):
            self.pos[i] = self.pos[i] + (self.second[i] * self.dt + self.forces[i] * term)
        self.clean_forces()
        self.calc_forces()
        for i in range(self.n):
            self.second[i] = self.second[i] + self.dt / 2 * (self.forces[i] + old[i])

    # src/integration.jl:433-465
    def update_szabo(self, step):
        d = self.dyn
        for i in range(self.n):
            theta = self.second[i]
            pol = np.array([math.cos(theta), math.sin(theta)])
            vel = d.vo * pol + d.mobility * self.forces[i]
            speed = math.sqrt(abs(vel[0]) + abs(vel[1]))
            cross = (pol[0] * vel[1] - pol[1] * vel[0]) / speed if speed > 0 else 0
            if abs(cross) > 1:
                cross = math.copysign(1.0, cross)
            d_theta = 1 / d.relax_time * math.asin(cross) * self.dt + math.sqrt(2 * d.rot_diff * self.dt) * self.noise[step][i]
            self.pos[i] = self.pos[i] + vel * self.dt
            self.second[i] = theta + d_theta

    # src/integration.jl:467-498 (u, u2 = the two rand() draws)
    def update_rtp(self, step):
        d = self.dyn
        for i in range(self.n):
            theta = self.second[i]
            pol = np.array([math.cos(theta), math.sin(theta)])
            vel = d.vo * pol + self.forces[i]
            self.pos[i] = self.pos[i] + vel * self.dt
            u, u2 = self.noise[step][i]
            if u < d.tumble_rate * self.dt:
                self.second[i] = 2 * math.pi * u2

    # src/integration.jl:271-285 / :309-324
    def walls(self):
        if self.periodic:
            center = self.bl + self.size / 2
            half = self.size / 2
            for i in range(self.n):
                diff = self.pos[i] - center
                out = np.abs(diff) > half
                if out.any():
                    self.pos[i] = self.pos[i] - np.sign(diff) * (half * 2) * out
        else:
            for i in range(self.n):
                rel = self.pos[i] - self.bl
                out = ((rel + self.radius) > self.size) | ((rel - self.radius) < 0)
                if out.any():
                    self.second[i] = self.second[i] * (-2 * out + 1)

    # src/integration.jl:507-535
    def step(self, nsteps):
        for s in range(nsteps):
            self.clean_forces()
            self.update_chunks()
            self.calc_forces()
            if isinstance(self.dyn, pkg.SzaboCfg):
                self.update_szabo(s)
            elif isinstance(self.dyn, pkg.RunTumbleCfg):
                self.update_rtp(s)
            else:
                self.update_verlet()
            self.walls()
            self.time += self.dt
            self.num_steps += 1


@pytest.mark.parametrize("dyn", ["lj", "harm"])
def test_newton_steps_match_second_restatement(oracle, dyn):
    d = pkg.LenJonesCfg(sigma=1.0, epsilon=1.0) if dyn == "lj" else pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2)
    # hot enough (vmax 3, dt 0.004) that particles change cell and wrap through the periodic walls within 12 steps
    case = H.newton_case(nx=9, ny=8, dyn=d, wall="periodic", jitter=0.3, vmax=3.0, dt=0.004, cells=(6, 5))
    o, r = H.make_oracle(case), RefSim(case)
    c0 = o.download_cells()[0].copy()
    o.step(12)
    r.step(12)
    assert np.abs(o.pos() - np.array(r.pos)).max() < 1e-13 * case["geom"].length
    assert H.rel_err(o.second(), np.array(r.second)) < 1e-12
    assert H.rel_err(o.get_forces(), np.array(r.forces)) < 1e-11
    assert o.time() == (r.num_steps, r.time)
    o.update_chunks()
    assert not np.array_equal(o.download_cells()[0], c0), "the case must re-bin particles to test stale-cell semantics"


@pytest.mark.parametrize("kind", ["szabo", "rtp"])
def test_self_propelled_steps_match_second_restatement(oracle, kind):
    case = H.sp_case(kind, nx=8, ny=7, jitter=0.9 if kind == "szabo" else 0.6, rot_diff=0.3)
    n, steps = 56, 10
    rng = np.random.default_rng(2)
    noise = rng.standard_normal((steps, n)) if kind == "szabo" else np.stack([rng.random((steps, n)) * 0.02, rng.random((steps, n))], -1)
    o, r = H.make_oracle(case), RefSim(case, noise=noise)
    o.step(steps, noise)
    r.step(steps)
    assert np.abs(o.pos() - np.array(r.pos)).max() < 1e-13 * case["geom"].length
    assert np.abs(o.second() - np.array(r.second)).max() < 1e-12
    assert H.rel_err(o.get_forces(), np.array(r.forces)) < 1e-11


def test_rigid_wall_newton_steps_match_second_restatement(oracle):
    """README quick start C1 in miniature: all-pairs LJ (src/integration.jl:197-224) with the velocity flips of the rigid
    rectangle's walls! (src/integration.jl:271-285) inside whole newton_step!s."""
    case = H.newton_case(nx=6, ny=5, wall="rigid", chunks=False, dt=0.01, jitter=0.2, vmax=2.0)

    # all-pairs restatement (src/integration.jl:197-224): i < j over the ids, Newton-3 scatter
    st = case["mk"]()
    pos = [p.copy() for p in st.pos]
    vel = [v.copy() for v in st.vel]
    n, dt = len(pos), case["int_cfg"].dt
    dyn = case["dyn"]
    size = np.array([case["geom"].length, case["geom"].height])
    radius = pkg.particle_radius(dyn)

    def forces_of(pos):
        F = [np.zeros(2) for _ in range(n)]
        for i in range(n):
            for j in range(i + 1, n):
                dr = pos[i] - pos[j]
                dist = math.sqrt(dr[0] ** 2 + dr[1] ** 2)
                fmod = 4 * dyn.epsilon * (12 * dyn.sigma ** 12 / dist ** 13 - 6 * dyn.sigma ** 6 / dist ** 7)
                f = fmod / dist * dr
                F[i] = F[i] + f
                F[j] = F[j] - f
        return F

    for _ in range(30):
        F = forces_of(pos)
        old = [f.copy() for f in F]
        for i in range(n):
            pos[i] = pos[i] + (vel[i] * dt + F[i] * (dt ** 2 / 2))
        F = forces_of(pos)
        for i in range(n):
            vel[i] = vel[i] + dt / 2 * (F[i] + old[i])
            out = ((pos[i] + radius) > size) | ((pos[i] - radius) < 0)
            if out.any():
                vel[i] = vel[i] * (-2 * out + 1)
    o = H.make_oracle(case)
    o.step(30)
    assert np.abs(o.pos() - np.array(pos)).max() < 1e-12 * case["geom"].length
    assert H.rel_err(o.second(), np.array(vel)) < 1e-11
    assert np.abs(np.array(vel) - st.vel).max() > 0.1
