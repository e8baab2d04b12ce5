"""Chain / world descriptions (host side, plain data).

These mirror what the reference reads from its ZTK model files ([roki::chain], [roki::link],
[roki::motor], [zeo::shape], [roki::contact]; reference example/model/*.ztk) and are the inputs
handed to the C-ABI (`rkFDChainReg` through the `rkChainB200*` builder calls).  Conventions:
revolute/prismatic axis = link-local z; `org_R`/`org_p` = the ZTK `frame:` of the link w.r.t. its
parent; inertia about the COM in link axes.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

NDOF = {"fixed": 0, "revolute": 1, "prismatic": 1, "spherical": 3, "float": 6, "cylindrical": 2, "hooke": 2, "breakablefloat": 6}


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], float)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], float)


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], float)


def box_verts(depth, width, height, center=(0.0, 0.0, 0.0)):
    """8 corners of a zeo box (depth=x, width=y, height=z); vertex k has sign bits (x: k&1, y: k&2, z: k&4)."""
    c = np.asarray(center, float)
    v = np.array([[(1 if k & 1 else -1) * depth / 2, (1 if k & 2 else -1) * width / 2,
                   (1 if k & 4 else -1) * height / 2] for k in range(8)], float)
    return v + c


@dataclass
class Motor:
    type: str = "dc"               # "dc" | "trq"
    k: float = 0.0                 # motorconstant
    admittance: float = 0.0
    gear: float = 1.0              # gearratio
    rotor_inertia: float = 0.0
    gear_inertia: float = 0.0
    min: float = -1e300
    max: float = 1e300


@dataclass
class Link:
    name: str = "link"
    parent: int = -1               # index within the chain, -1 = root
    jtype: str = "fixed"
    org_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    org_p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    mass: float = 1.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.eye(3))
    stuff: str = ""
    stiffness: float = 0.0
    viscosity: float = 0.0
    coulomb: float = 0.0
    sfriction: float = 0.0
    motor: Optional[Motor] = None
    break_force: float = 0.0       # breakablefloat: force / torque thresholds (ZTK `forcethreshold` / `torquethreshold`)
    break_torque: float = 0.0
    shapes: List[np.ndarray] = field(default_factory=list)   # vertex clouds (n x 3, link frame)
    # box primitives as (center(3), depth, width, height) in the link frame.  On a static link: a collision target.  On a
    # moving link: a target for the vertices of OTHER links' cells, and its 8 corners are a cell of this link (after `shapes`)
    boxes: List[tuple] = field(default_factory=list)
    # slide mode ("fake crawler", rkFDCDCellSetSlideMode/-Vel/-Axis, rkfd_sim.c:386-440) per collision cell of the link:
    # {cell index in the link (shapes first, then boxes): (belt speed, axis(3) in the link frame)}
    slides: dict = field(default_factory=dict)

    def cells(self):
        """Vertex clouds of a moving link in cell order: the shapes, then the corners of its box primitives."""
        return list(self.shapes) + [box_verts(d, w, h, center=c) for (c, d, w, h) in self.boxes]


@dataclass
class ChainModel:
    name: str
    links: List[Link]
    # pairs between the cells of this chain itself: the reference registers them until rkCDPairChainUnreg (every example
    # program calls it); worlds built here default to the unregistered state
    self_collide: bool = False

    @property
    def joint_size(self):
        return sum(NDOF[l.jtype] for l in self.links)

    @property
    def is_static(self):
        return self.joint_size == 0


@dataclass
class StaticBox:
    R: np.ndarray
    p: np.ndarray
    half: np.ndarray
    stuff: str
    slide: Optional[tuple] = None          # (belt speed, axis(3) in the frame of the box's link)
    link_R: np.ndarray = field(default_factory=lambda: np.eye(3))     # world frame of the (static) link that carries the box
    link_p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    order: int = 0                         # registration order of the shape among all shapes of the world


@dataclass
class ContactInfo:
    stuff_a: str
    stuff_b: str
    type: str = "rigid"            # "rigid" (compensation/relaxation) | "elastic" (elasticity/viscosity)
    K: float = 0.0
    L: float = 0.0
    E: float = 0.0
    V: float = 0.0
    SF: float = 0.5
    KF: float = 0.3


@dataclass
class World:
    """One environment's content: chains registered in order + contact table + properties
    (reference rkfd_property.c:10-18 defaults)."""
    chains: List[ChainModel] = field(default_factory=list)
    contact_info: List[ContactInfo] = field(default_factory=list)
    dt: float = 0.001
    pyramid: int = 8
    friction_weight: float = 100.0
    max_iter: int = 10
    solver: str = "Vert"
    integrator: str = "RKG"        # zODE2AssignRegular: "RKG" (reference default) | "RK4" | "Euler" | "Heun"

    def __post_init__(self):
        self._stuff = {}

    def stuff_id(self, name):
        if not hasattr(self, "_stuff"):
            self._stuff = {}
        return self._stuff.setdefault(name, len(self._stuff))

    def moving_chains(self):
        return [c for c in self.chains if not c.is_static]

    def flat_links(self):
        """Links of all moving chains concatenated in registration order, parents re-indexed."""
        out = []
        for ch in self.moving_chains():
            base = len(out)
            for l in ch.links:
                ll = Link(**{**l.__dict__})
                ll.parent = l.parent + base if l.parent >= 0 else -1
                out.append(ll)
        return out

    @property
    def boxes(self):
        """Static boxes in world coordinates from the all-fixed chains (e.g. floor.ztk)."""
        out = []
        for ch in self.chains:
            if not ch.is_static:
                continue
            frames = []
            for l in ch.links:
                R, p = np.asarray(l.org_R, float), np.asarray(l.org_p, float)
                if l.parent >= 0:
                    Rp, pp = frames[l.parent]
                    R, p = Rp @ R, pp + Rp @ p
                frames.append((R, p))
                for bi, (c, d, w, h) in enumerate(l.boxes):
                    out.append(StaticBox(R=R, p=p + R @ np.asarray(c, float),
                                         half=np.array([d / 2, w / 2, h / 2], float), stuff=l.stuff,
                                         slide=l.slides.get(bi), link_R=R, link_p=p,
                                         order=self.shape_order()[(id(ch), id(l), bi)]))
        return out

    def shape_order(self):
        """Registration order of every collision shape: chains in the order of rkFDChainReg, links, shapes then boxes.  Decides
        which cell of a pair is pd->cell[0] (rkFDUpdateRefSlide, rkfd_util.c:218-237)."""
        out, n = {}, 0
        for ch in self.chains:
            for l in ch.links:
                ncell = len(l.boxes) if ch.is_static else len(l.shapes) + len(l.boxes)
                for k in range(ncell):
                    out[(id(ch), id(l), k)] = n; n += 1
        return out

    @property
    def nq(self):
        return sum(c.joint_size for c in self.moving_chains())

    @property
    def nl(self):
        return sum(len(c.links) for c in self.moving_chains())

    def _ci_type(self, sa, sb):
        for ci in self.contact_info:
            if (ci.stuff_a, ci.stuff_b) in ((sa, sb), (sb, sa)):
                return ci.type
        return "rigid"                     # every solver's default contact info is rigid

    @property
    def nslot(self):
        """Contact slots = sum over pairs of the cell's vertices: every cell of a moving link against every static box, and
        against the box primitives of other moving links (other chains; the same chain only with `self_collide`; under the
        Volume solver elastic contact info only) - the rule of rkfd_model.cpp / the oracle's ork_world_finalize."""
        links = self.flat_links()
        n = sum(v.shape[0] for l in links for v in l.cells()) * len(self.boxes)
        chain_of, self_col = [], []
        for ci, chn in enumerate(self.moving_chains()):
            chain_of += [ci] * len(chn.links); self_col.append(chn.self_collide)
        for a, la in enumerate(links):
            for b, lb in enumerate(links):
                if a == b or not lb.boxes or (chain_of[a] == chain_of[b] and not self_col[chain_of[a]]):
                    continue
                if self._ci_type(la.stuff, lb.stuff) != "elastic" and self.solver == "Volume":
                    continue
                n += sum(v.shape[0] for v in la.cells()) * len(lb.boxes)
        return n


def world_from_flat(desc):
    """A World equivalent to a flattened device model (`RkFD.describe_model()`: what `rkFDUpdateInit` made of the
    registered chains - e.g. of the reference's own ZTK files read through `rkFDChainRegFile`).  Every moving link goes
    into one forest chain, every static box into its own all-fixed chain; the contact parameters of each (cell, box) pair
    come back through per-link / per-box `stuff` names.  Flattening the result again reproduces the description (the DC
    motor constants are folded in the description: an equivalent motor with gear ratio 1 is returned)."""
    jt = {0: "fixed", 1: "revolute", 2: "prismatic", 3: "spherical", 4: "float", 5: "cylindrical", 6: "hooke", 7: "breakablefloat"}
    nl, nq, ncell, nbox, npair, nslot, nvert = (int(v) for v in desc["dims"])
    solver, pyramid, max_iter, integ = (int(v) for v in desc["prp"][:4])
    dt, fw = desc["prp"][4], desc["prp"][5]
    verts = np.array([desc["vert[%d]" % i] for i in range(nvert)], float).reshape(nvert, 3)
    cells = [[int(v) for v in desc["cell[%d]" % i]] for i in range(ncell)]
    links = []
    for i in range(nl):
        topo = [int(v) for v in desc["link.topo[%d]" % i]]
        mp, jf = desc["link.mass[%d]" % i], desc["link.joint[%d]" % i]
        m, com = mp[0], np.array(mp[10:13])
        Io = np.array([[mp[4], mp[5], mp[6]], [mp[5], mp[7], mp[8]], [mp[6], mp[8], mp[9]]])
        Ic = Io - m * (com @ com * np.eye(3) - np.outer(com, com))
        l = Link(name="link#%02d" % i, parent=topo[0], jtype=jt[topo[1]], org_R=np.array(desc["link.Ro[%d]" % i]).reshape(3, 3),
                 org_p=np.array(desc["link.po[%d]" % i]), mass=m, com=com, inertia=Ic, stuff="L%d" % i,
                 stiffness=jf[0], viscosity=jf[1], coulomb=jf[2], sfriction=jf[3])
        if topo[2] == 1:        # DC motor: m_tin = g k a, m_reg = (g k)^2 a, m_jm = g^2 (Jr + Jg)  ->  g = 1
            k = jf[5] / jf[4] if jf[4] != 0.0 else 0.0
            a = jf[4] * jf[4] / jf[5] if jf[5] != 0.0 else 0.0
            l.motor = Motor(type="dc", k=k, admittance=a, gear=1.0, rotor_inertia=jf[6], gear_inertia=0.0, min=jf[7], max=jf[8])
        elif topo[2] == 2:
            l.motor = Motor(type="trq", min=jf[7], max=jf[8])
        l.shapes = [verts[c[1]:c[1] + c[2]].copy() for c in cells[topo[5]:topo[6]]]
        links.append(l)
    chains = [ChainModel("flat", links)] if links else []
    for b in range(nbox):
        h = desc["box.half[%d]" % b]
        chains.append(ChainModel("box%d" % b, [Link(name="box", jtype="fixed", stuff="B%d" % b, org_R=np.array(desc["box.R[%d]" % b]).reshape(3, 3),
                                                     org_p=np.array(desc["box.p[%d]" % b]), boxes=[((0.0, 0.0, 0.0), 2 * h[0], 2 * h[1], 2 * h[2])])]))
    # slide mode of moving cells (rkFDShape3DSetSlide* before rkFDUpdateInit): back onto the links' cells
    for p in range(npair):
        sl = desc.get("pair.slide[%d]" % p)
        if sl is None:
            continue
        if int(sl[1]) > 0:
            raise NotImplementedError("world_from_flat: a static box in slide mode (its link frame is not part of the description)")
        c = int(desc["pair[%d]" % p][0]); e = desc["slide[%d]" % (int(sl[0]) - 1)]
        lk = cells[c][0]; first = int(desc["link.topo[%d]" % lk][5])
        links[lk].slides[c - first] = (e[0], tuple(e[1:4]))
    ci, seen = [], set()
    for p in range(npair):
        v = desc["pair[%d]" % p]
        key = ("L%d" % cells[int(v[0])][0], "B%d" % int(v[1]))
        if key in seen:
            continue
        seen.add(key)
        ci.append(ContactInfo(key[0], key[1], "rigid" if int(v[3]) == 0 else "elastic", K=v[4], L=v[5], E=v[6], V=v[7], SF=v[8], KF=v[9]))
    return World(chains=chains, contact_info=ci, dt=dt, pyramid=pyramid, friction_weight=fw, max_iter=max_iter,
                 solver={0: "Vert", 1: "MLCP", 2: "Volume"}[solver], integrator={0: "RKG", 1: "RK4", 2: "Euler", 3: "Heun"}[integ])


def load_flat(path):
    """A description saved as text (one `name: numbers` line per table entry) -> dict."""
    out = {}
    with open(path) as f:
        for line in f:
            k, _, v = line.partition(":")
            if v.strip():
                out[k] = [float(x) for x in v.split()]
    return out


# ---------------------------------------------------------------------------------------------
# models of the reference's example/model directory (values transcribed from the ZTK files)

def motor_arm2dof():
    """[roki::motor] motor1 of reference example/model/arm_2DoF.ztk:131-140."""
    return Motor(type="dc", k=2.58e-2, admittance=0.42373, gear=120.0, rotor_inertia=1.65e-6,
                 gear_inertia=5.38e-6, min=-24.0, max=24.0)


def box(name="box"):
    """reference example/model/box.ztk: one float link, m=0.5, 0.1 m cube, stuff body."""
    return ChainModel(name, [Link(name="link#00", jtype="float", mass=0.5, stuff="body",
                                  inertia=np.eye(3) * 8.33e-4, shapes=[box_verts(0.1, 0.1, 0.1)])])


def floor():
    """reference example/model/floor.ztk: fixed link, 5 x 5 x 0.4 box, top face at z=0, stuff ground."""
    return ChainModel("floor", [Link(name="link#00", jtype="fixed", mass=99.9, stuff="ground",
                                     inertia=np.eye(3) * 0.999, boxes=[((0, 0, -0.2), 5.0, 5.0, 0.4)])])


def floor_soft():
    """floor.ztk with `stuff: soft`."""
    f = floor()
    f.links[0].stuff = "soft"
    return f


def floor_hardsoft():
    """reference example/model/floor_hardsoft.ztk:31-75: two 5 x 2.5 x 0.4 boxes, y>0 `ground`, y<0 `soft`."""
    l0 = Link(name="link#00", jtype="fixed", mass=99.9, stuff="ground", inertia=np.eye(3) * 0.999,
              org_p=np.array([0, 1.25, 0.0]), boxes=[((0, 0, -0.2), 5.0, 2.5, 0.4)])
    l1 = Link(name="link#01", jtype="fixed", parent=0, mass=99.9, stuff="soft", inertia=np.eye(3) * 0.999,
              org_p=np.array([0, -2.5, 0.0]), boxes=[((0, 0, -0.2), 5.0, 2.5, 0.4)])
    return ChainModel("floor", [l0, l1])


def contact_info_table():
    """reference example/model/contactinfo.ztk (all seven records)."""
    return [
        ContactInfo("ground", "body", "rigid", K=1000.0, L=0.0001, SF=0.5, KF=0.3),
        ContactInfo("body", "body", "rigid", K=1000.0, L=0.05, SF=0.5, KF=0.3),
        ContactInfo("wall", "wall", "rigid", K=500.0, L=0.001, SF=0.5, KF=0.3),
        ContactInfo("wall", "ground", "rigid", K=500.0, L=0.001, SF=0.5, KF=0.3),
        ContactInfo("ground", "crawler", "rigid", K=100.0, L=10.0, SF=10.0, KF=7.0),
        ContactInfo("soft", "body", "elastic", E=100.0, V=1.0, SF=0.5, KF=0.3),
        ContactInfo("soft", "crawler", "elastic", E=1000.0, V=10.0, SF=10.0, KF=7.0),
    ]


def arm_2dof(motors=True):
    """reference example/model/arm_2DoF.ztk:142-212: fixed base + 2 revolute links with DC motors."""
    m = motor_arm2dof if motors else (lambda: None)
    base = Link(name="link#00", jtype="fixed", mass=1.5, stuff="body", com=np.array([0.067, 0, 0]),
                inertia=np.diag([7.395833e-03, 8.645833e-03, 8.541667e-03]),
                org_R=np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0]], float))
    l1 = Link(name="link#01", jtype="revolute", parent=0, mass=1.5, stuff="body", com=np.array([0.267, 0, 0]),
              inertia=np.diag([0.00239583, 0.02239583, 0.02229167]), org_p=np.array([0.15, 0, 0]),
              viscosity=2.2, coulomb=4.32, sfriction=4.92, motor=m())
    l2 = Link(name="link#02", jtype="revolute", parent=1, mass=1.0, stuff="body", com=np.array([0.2, 0, 0]),
              inertia=np.diag([0.0016667, 0.0083333, 0.0083333]), org_p=np.array([0.4, 0, 0]),
              viscosity=2.2, coulomb=4.32, sfriction=4.92, motor=m())
    return ChainModel("arm_2DoF", [base, l1, l2])


# ---------------------------------------------------------------------------------------------
# synthetic models of BASELINE.json / SURVEY.md section 8d

ARM7_MASS = [4.0, 4.0, 3.0, 2.7, 1.7, 1.8, 0.3]


def arm7(base_z=0.6, contact_cube=False, motors=True, stuff="body"):
    """The 7-DoF arm of BASELINE.json configs C2/C3/C5: fixed base + 7 revolute
    links, joint axis local z, link frame (0,0,0.2) + Rot_x(-/+90 deg) alternating, DC motor + joint
    friction constants of arm_2DoF.ztk on every joint.  `contact_cube` hangs the 0.1 m cube of box.ztk
    (8 vertices) on link 7."""
    links = [Link(name="base", jtype="fixed", mass=5.0, stuff=stuff, inertia=np.eye(3) * 0.05,
                  org_p=np.array([0, 0, base_z]))]
    for i in range(1, 8):
        m = ARM7_MASS[i - 1]
        l = Link(name="link#%02d" % i, jtype="revolute", parent=i - 1, mass=m, stuff=stuff,
                 com=np.array([0, 0, 0.1]), inertia=np.diag([m * 0.01, m * 0.01, m * 0.002]),
                 org_p=np.array([0, 0, 0.2]),
                 org_R=rot_x(-np.pi / 2) if i % 2 else rot_x(np.pi / 2))
        if motors:
            l.motor = motor_arm2dof()
            l.viscosity, l.coulomb, l.sfriction = 2.2, 4.32, 4.92
        links.append(l)
    if contact_cube:
        links[7].shapes = [box_verts(0.1, 0.1, 0.1, center=(0, 0, 0.1))]
    return ChainModel("arm7", links)


def world_c2():
    """C2: arm7, no contact (pure ABA + joint friction)."""
    return World(chains=[arm7()])


def world_c3(base_z=0.45):
    """C3: arm7 + end-effector cube on a soft (elastic E=1000, V=10, SF=.5, KF=.3) floor, Vert solver."""
    return World(chains=[arm7(base_z=base_z, contact_cube=True), floor_soft()],
                 contact_info=[ContactInfo("soft", "body", "elastic", E=1000.0, V=10.0, SF=0.5, KF=0.3)])


def world_c5(base_z=0.45, solver="MLCP"):
    """C5: arm7 + cube on the rigid `ground` floor (K=1000, L=1e-4), MLCP or Vert-QP."""
    return World(chains=[arm7(base_z=base_z, contact_cube=True), floor()],
                 contact_info=[ContactInfo("ground", "body", "rigid", K=1000.0, L=0.0001, SF=0.5, KF=0.3)],
                 solver=solver)


def world_c1_box(solver="Vert"):
    """C1: box.ztk dropped on floor_hardsoft.ztk with contactinfo.ztk (boxdrop_hardsoft_test.c)."""
    return World(chains=[box(), floor_hardsoft()], contact_info=contact_info_table(), solver=solver)


def world_c1_serial():
    """C1 variant: arm_2DoF.ztk falling under gravity with motors at 0 V, no ground."""
    return World(chains=[arm_2dof()])


def biped(stuff="body"):
    """A small legged tree for the C4-shaped workload (multi-limb legged model): floating
    trunk + two legs of three revolute joints (hip, knee, ankle: pitch axes) with DC motors and an 8-vertex sole
    under each ankle link.  12 DoF, 7 links, 16 contact vertices.  (The reference's humanoid mighty.ztk has 26 DoF
    and polyhedral soles; this tree keeps box soles, the shape the Volume solver here covers.)"""
    links = [Link(name="trunk", jtype="float", mass=8.0, stuff=stuff, com=np.array([0, 0, 0.1]),
                  inertia=np.diag([0.12, 0.10, 0.06]))]
    for side, y in (("L", 0.09), ("R", -0.09)):
        base = len(links)
        specs = [("hip", 0, np.array([0.0, y, -0.05]), 1.5, 0.18), ("knee", base, np.array([0.0, -0.2, 0.0]), 1.0, 0.18),
                 ("ankle", base + 1, np.array([0.0, -0.2, 0.0]), 0.4, 0.03)]
        for k, (nm, parent, p, m, ln) in enumerate(specs):
            l = Link(name=nm + side, parent=parent, jtype="revolute", mass=m, stuff=stuff,
                     com=np.array([0.0, -ln / 2, 0.0]), inertia=np.diag([m * 0.004, m * 0.001, m * 0.004]),
                     org_p=p, org_R=rot_x(np.pi / 2) if k == 0 else np.eye(3))
            l.motor = motor_arm2dof()
            l.viscosity, l.coulomb, l.sfriction = 2.2, 4.32, 4.92
            links.append(l)
        links[-1].shapes = [box_verts(0.16, 0.03, 0.08, center=(0.03, -0.045, 0.0))]
    return ChainModel("biped", links)


def world_c4_penalty():
    """C4-shaped workload with what is built: the legged tree on the soft floor (vertex penalty contact)."""
    return World(chains=[biped(), floor_soft()],
                 contact_info=[ContactInfo("soft", "body", "elastic", E=1000.0, V=10.0, SF=0.5, KF=0.3)])


def world_c4_volume():
    """C4 (BASELINE.json configs[3]): the legged tree on the rigid floor with volume-based contact (rkfd_volume),
    the solver's default contact info (rkfd_volume.c:961-969)."""
    return World(chains=[biped(), floor()], solver="Volume")


def sample_c4_standing(world, B, seed=20260418):
    """C4 workload states: the legged tree standing on both soles (trunk 0.512 m +- 2 mm above the floor, joint angles
    +- 0.05 rad, velocities +- 0.05, motors off: joint friction holds the pose), i.e. both contact volumes live in
    every environment and every step."""
    rng = np.random.default_rng(seed)
    q = np.zeros((B, world.nq)); qd = rng.uniform(-0.05, 0.05, (B, world.nq)); u = np.zeros((B, world.nl))
    q[:, 2] = 0.512 + rng.uniform(-0.002, 0.002, B)
    q[:, 3:6] = rng.uniform(-0.02, 0.02, (B, 3))
    q[:, 6:] = rng.uniform(-0.05, 0.05, (B, world.nq - 6))
    return q, qd, u


def sample_state(world, B, seed=20260418):
    """Synthetic randomised states of SURVEY.md section 8d: q ~ U(-pi/2, pi/2), qd ~ U(-1, 1),
    motor voltage ~ U(-6, 6); float joints get position z lifted by +0.3."""
    rng = np.random.default_rng(seed)
    nq, nl = world.nq, world.nl
    q = rng.uniform(-np.pi / 2, np.pi / 2, (B, nq))
    qd = rng.uniform(-1.0, 1.0, (B, nq))
    u = rng.uniform(-6.0, 6.0, (B, nl))
    o = 0
    for l in world.flat_links():
        if l.jtype == "float":
            q[:, o:o + 3] = rng.uniform(-0.2, 0.2, (B, 3))
            q[:, o + 2] += 0.3
        o += NDOF[l.jtype]
    return q, qd, u


# ---------------------------------------------------------------------------------------------
# random chains for property tests

def random_chain(rng, n_links=5, jtypes=("revolute",), root="fixed", branching=False, motors=False,
                 shapes=False, name="rnd"):
    def rnd_rot():
        a = rng.normal(size=3)
        th = np.linalg.norm(a)
        k = a / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K

    def rnd_inertia(m):
        A = rng.normal(size=(3, 3))
        return (A @ A.T + 0.5 * np.eye(3)) * 0.01 * m

    links = []
    for i in range(n_links):
        m = rng.uniform(0.5, 3.0)
        jt = root if i == 0 else jtypes[rng.integers(len(jtypes))]
        parent = -1 if i == 0 else (int(rng.integers(0, i)) if branching else i - 1)
        l = Link(name="l%d" % i, parent=parent, jtype=jt, mass=m, com=rng.uniform(-0.1, 0.1, 3),
                 inertia=rnd_inertia(m), org_R=rnd_rot(), org_p=rng.uniform(-0.3, 0.3, 3), stuff="body")
        if i == 0 and root == "fixed":
            l.org_p = np.array([0, 0, 1.0])
        if motors and NDOF[jt] == 1:
            l.motor = motor_arm2dof()
            l.viscosity, l.coulomb, l.sfriction = 2.2, 4.32, 4.92
        if shapes and i == n_links - 1:
            l.shapes = [box_verts(0.1, 0.1, 0.1)]
        links.append(l)
    return ChainModel(name, links)
