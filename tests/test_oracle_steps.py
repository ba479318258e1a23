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


# ---------------------------------------------------------------- Mavi.Rings: whole steps vs a second restatement
class RefRings:
    """Plain-Python restatement of the Rings constructor tail and step! (src/rings/rings.jl:276-288,
    src/rings/integration.jl:32-226, :300-372, :522-543), all-pairs pair enumeration (src/integration.jl:197-224)."""

    def __init__(self, case, noise):
        st = case["mk"]()
        self.nr, self.nmax = st.num_rings, st.n_max
        self.pos = [np.array(p, dtype=np.float64) for p in st.pos]          # scalar idx = ring * n_max + p
        self.pol = [float(a) for a in st.pol]
        self.types = None if st.types is None else [int(t) - 1 for t in st.types]
        self.dyn = case["dyn"]
        self.np_of = [st.ring_num_particles(r) for r in range(self.nr)]
        self.size = np.array([case["geom"].length, case["geom"].height])
        self.periodic = isinstance(case["space"].wall_type, pkg.PeriodicWalls)
        self.dt = case["int_cfg"].dt
        self.noise = noise
        n = len(self.pos)
        self.cont = [np.zeros(2) for _ in range(n)]
        self.cms = [np.zeros(2) for _ in range(self.nr)]
        self.areas = [0.0] * self.nr
        self.forces = [np.zeros(2) for _ in range(n)]
        self.ids = [r * self.nmax + q for r in range(self.nr) for q in range(self.np_of[r])]
        # RingsSystem ctor tail, src/rings/rings.jl:280-288
        self.update_continuos_pos()
        self.update_cms()
        self.clean()
        self.forces_()

    def ty(self, ring):
        return 0 if self.types is None else self.types[ring]

    def calc_diff(self, r1, r2):
        dr = r1 - r2
        if self.periodic:
            dr = dr - (np.abs(dr) > (self.size / 2)) * np.copysign(self.size, dr)
        return dr

    def ring_points(self, ring):  # get_continuos_pos, src/rings/rings.jl:31-43
        src = self.cont if self.periodic else self.pos
        return [src[ring * self.nmax + q] for q in range(self.np_of[ring])]

    def update_continuos_pos(self):  # src/rings/integration.jl:118-138
        if not self.periodic:
            return
        for ring in range(self.nr):
            b = ring * self.nmax
            for q in range(self.nmax):
                self.cont[b + q] = self.pos[b + q].copy()
            for q in range(1, self.np_of[ring]):
                self.cont[b + q] = self.cont[b + q - 1] + self.calc_diff(self.pos[b + q], self.pos[b + q - 1])

    def update_cms(self):  # src/rings/integration.jl:366-372
        for ring in range(self.nr):
            pts = self.ring_points(ring)
            self.cms[ring] = sum(pts[1:], pts[0].copy()) / len(pts)

    def clean(self):
        self.forces = [np.zeros(2) for _ in self.pos]

    def calc_interaction(self, i, j):  # src/rings/integration.jl:32-77
        ri, rj = i // self.nmax, j // self.nmax
        ic = self.dyn.interaction(self.ty(ri), self.ty(rj))
        dr = self.calc_diff(self.pos[i], self.pos[j])
        dist = math.sqrt(dr[0] ** 2 + dr[1] ** 2)
        if dist > ic.dist_max:
            return np.zeros(2)
        if ri == rj:
            diff = abs(i - j)
            if diff == 1 or diff == self.np_of[ri] - 1:
                return np.zeros(2)
        if dist < ic.dist_eq:
            fmod = -ic.k_rep * (dist / ic.dist_eq - 1)
        elif ri == rj:
            fmod = 0.0
        else:
            fmod = -ic.k_atr * (dist / ic.dist_eq - 1)
        return fmod / dist * dr

    def forces_(self):  # forces!, src/rings/integration.jl:197-226
        ids = self.ids
        for a in range(len(ids)):
            for b in range(a + 1, len(ids)):
                f = self.calc_interaction(ids[a], ids[b])
                self.forces[ids[a]] = self.forces[ids[a]] + f
                self.forces[ids[b]] = self.forces[ids[b]] - f
        d = self.dyn
        for ring in range(self.nr):
            n, t, b = self.np_of[ring], self.ty(ring), ring * self.nmax
            k, l = d.k_spring[t], d.l_spring[t]
            for s in range(n):   # springs_force, :79-97
                p1, p2 = b + s, b + (0 if s == n - 1 else s + 1)
                dr = self.calc_diff(self.pos[p1], self.pos[p2])
                dist = math.sqrt(dr[0] ** 2 + dr[1] ** 2)
                f = (-k * (dist - l)) / dist * dr
                self.forces[p1] = self.forces[p1] + f
                self.forces[p2] = self.forces[p2] - f
        for ring in range(self.nr):  # area_forces!, :140-195 (+ calc_area :103-116)
            n, t, b = self.np_of[ring], self.ty(ring), ring * self.nmax
            pts = self.ring_points(ring)
            area = 0.0
            for q in range(n - 1):
                area += pts[q][0] * pts[q + 1][1] - pts[q][1] * pts[q + 1][0]
            area += pts[-1][0] * pts[0][1] - pts[-1][1] * pts[0][0]
            area = area / 2.0
            self.areas[ring] = area
            area0 = (n * d.l_spring[t] / d.p0[t]) ** 2
            for q in range(n):
                fmod = d.k_area[t] * (area - area0)
                id1 = n - 1 if q == 0 else q - 1
                id2 = 0 if q == n - 1 else q + 1
                dr = self.calc_diff(self.pos[b + id2], self.pos[b + id1])
                a_deriv = np.array([dr[1], -dr[0]]) / 2
                self.forces[b + q] = self.forces[b + q] - fmod * a_deriv

    def update(self, step):  # update!, src/rings/integration.jl:300-351
        d = self.dyn
        for ring in range(self.nr):
            n, t, b = self.np_of[ring], self.ty(ring), ring * self.nmax
            theta = self.pol[ring]
            pol = np.array([math.cos(theta), math.sin(theta)])
            vel_cm = np.zeros(2)
            for q in range(n):
                vel = d.vo[t] * pol + d.mobility[t] * self.forces[b + q]
                vel_cm = vel_cm + vel
                self.pos[b + q] = self.pos[b + q] + vel * self.dt
            vel_cm = vel_cm / n
            speed = math.sqrt(vel_cm[0] ** 2 + vel_cm[1] ** 2)
            if speed == 0:
                cross = 0
            else:
                cross = (pol[0] * vel_cm[1] - pol[1] * vel_cm[0]) / speed
                if abs(cross) > 1:
                    cross = math.copysign(1.0, cross)
            self.pol[ring] = theta + (1 / d.relax_time[t] * math.asin(cross) * self.dt +
                                      math.sqrt(2 * d.rot_diff[t] * self.dt) * self.noise[step][ring])

    def walls(self):  # the core periodic walls! over the active ids (src/integration.jl:309-324)
        if not self.periodic:
            return
        half = self.size / 2
        for i in self.ids:
            diff = self.pos[i] - half
            out = np.abs(diff) > half
            if out.any():
                self.pos[i] = self.pos[i] - np.sign(diff) * (half * 2) * out

    def step(self, nsteps):  # step!, src/rings/integration.jl:522-543
        for s in range(nsteps):
            self.update_cms()
            self.update_continuos_pos()
            self.clean()
            self.forces_()
            self.update(s)
            self.walls()


@pytest.mark.parametrize("kind,use_chunks", [("normal", False), ("normal", True), ("types", True)])
def test_rings_steps_match_second_restatement(oracle, kind, use_chunks):
    n = 4
    case = H.rings_case(kind, n, n, use_chunks=use_chunks)
    steps = 40
    noise = np.random.default_rng(9).standard_normal((steps, n * n))
    o, r = H.make_oracle(case), RefRings(case, noise)
    assert H.rel_err(o.get_forces(), np.array(r.forces)) < 1e-12      # constructor state
    o.step(steps, noise)
    r.step(steps)
    assert np.abs(o.pos() - np.array(r.pos)).max() < 1e-12 * case["geom"].length
    assert np.abs(o.second() - np.array(r.pol)).max() < 1e-11
    assert H.rel_err(o.get_forces(), np.array(r.forces)) < 1e-10
    areas, cms, cont = o.rings_info()
    assert H.rel_err(areas, np.array(r.areas)) < 1e-12 and H.rel_err(cms, np.array(r.cms)) < 1e-12
    active = np.array(r.ids)
    assert H.rel_err(cont[active], np.array(r.cont)[active]) < 1e-12
