"""compute_parallel_aats -- drop-in for apyib/parallel.py:16-48.

The reference forks a 4-process multiprocessing.Pool over the (alpha, beta) tensor elements and
solves all finite-difference points serially first.  Here the unit of parallelism is the GPU:

  phase 1  the 6N+6 displaced / field points are partitioned over the ranks (one process per
           GPU, torch.distributed); every rank runs full, independent solves (no collective);
  exchange one all_gather of the per-point (C, T_list) payloads as raw device bytes (NCCL over
           NVLink; metadata only is pickled) -- every AAT element needs the amplitudes of the
           unperturbed, R+-alpha and B+-beta points;
  phase 2  tensor rows alpha are partitioned over the ranks; each rank evaluates its rows with
           the fused determinant kernels;
  gather   final all-gather of the (3N, 3) float64 tensor (NCCL on GPUs, gloo in CPU tests).

With world_size == 1 (or no process group) it degenerates to the single-GPU path.
`num_processes` is accepted for signature compatibility and ignored.
"""
from __future__ import annotations

import numpy as np

from . import config
from .aats import AAT
from .fin_diff import finite_difference, aat_points, point_cost
from .hostchem import Hamiltonian, hf_wfn


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1


def partition(items, costs, world):
    """Longest-processing-time static partition; deterministic, identical on every rank."""
    order = sorted(range(len(items)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owner = [0] * len(items)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def exchange_points(dist, blob, world):
    """The one exchange step of the path: every rank contributes {point: payload} for the points it
    solved and receives the union (every AAT element needs all amplitudes, aats.py:690-711).

    payload = (nested) list / tuple of numpy arrays, torch tensors and python scalars.  The array bytes
    of a rank travel as ONE padded uint8 buffer through a single all_gather -- NCCL over NVLink between
    device buffers on GPUs (amplitudes that are already device-resident never touch the host), gloo in
    the CPU tests; only the few hundred bytes of metadata (point, dtypes, shapes, scalars) are pickled.
    Arrays come back as device tensors under NCCL and as numpy arrays under gloo."""
    if dist is None or world == 1:
        return dict(blob)
    import torch
    nccl = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    meta, chunks, nbytes = [], [], 0

    def pack(x):
        nonlocal nbytes
        if isinstance(x, (list, tuple)):
            return ("seq", type(x).__name__, [pack(y) for y in x])
        if isinstance(x, np.ndarray) or torch.is_tensor(x):
            t = (torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x.detach().contiguous()).to(dev)
            raw = t.reshape(-1).view(torch.uint8)
            pad = (-raw.numel()) % 16                      # keep every slice 16-byte aligned (complex128 views)
            chunks.append(raw)
            if pad:
                chunks.append(torch.zeros(pad, dtype=torch.uint8, device=dev))
            m = ("arr", str(t.dtype).replace("torch.", ""), tuple(t.shape), nbytes, raw.numel())
            nbytes += raw.numel() + pad
            return m
        return ("obj", x)

    for pt in sorted(blob):
        meta.append((pt, pack(blob[pt])))
    metas = [None] * world
    dist.all_gather_object(metas, (meta, nbytes))
    width = max(16, max(m[1] for m in metas))
    buf = torch.zeros(width, dtype=torch.uint8, device=dev)
    if chunks:
        buf[:nbytes] = torch.cat(chunks)
    parts = [torch.empty(width, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)

    def unpack(m, raw):
        if m[0] == "seq":
            seq = [unpack(y, raw) for y in m[2]]
            return tuple(seq) if m[1] == "tuple" else seq
        if m[0] == "arr":
            _, dt, shape, off, n = m
            t = raw[off:off + n].view(getattr(torch, dt)).reshape(shape)
            return t if nccl else t.numpy()
        return m[1]

    out = {}
    for r in range(world):
        for pt, m in metas[r][0]:
            out[pt] = blob[pt] if pt in blob else unpack(m, parts[r])
    return out


def owned_elements(n3, rank, world):
    """(alpha, beta) elements of the (3N,3) tensor evaluated by `rank`: whole rows alpha, round
    robin.  All determinant families that depend on alpha (pu/nu[alpha], pp/pn/np/nn[alpha][:]) are
    then private to one rank (AAT(..., rows=) builds overlaps, norms and scaled amplitudes for the owned rows
    only); the 7 alpha-independent overlaps (uu, up/un) are evaluated per rank, against the owned bra rows."""
    return [(a, b) for a in range(n3) if a % world == rank for b in range(3)]


def gather_tensor(dist, I, world):
    """Final gather of the (3N,3) tensor: ranks fill disjoint elements, the rest is zero, so a
    sum-all-reduce is the gather (NCCL on GPUs, gloo in the CPU tests)."""
    if dist is None or world == 1:
        return I
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(I)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()


def compute_parallel_aats(parameters, nuc_pert_strength, mag_pert_strength, normalization='full', num_processes=4):
    from .energy import scf_point
    dist, rank, world = _dist()
    # every rank needs the unperturbed orbitals (phase reference of all points): host SCF on every rank; its
    # correlated solve is ONE more point of the partition, solved by one rank and exchanged with the others
    wfn = scf_point(parameters)
    C, basis = wfn.C, wfn.H.basis_set
    natom = wfn.H.molecule.natom()

    fd = finite_difference(parameters, basis, C)
    pts = [("U", 0, 0)] + aat_points(natom)
    owner = partition(pts, [point_cost(p[0]) for p in pts], world)
    mine = [p for p, o in zip(pts, owner) if o == rank]
    own_u = ("U", 0, 0) in mine
    lists = fd.compute_AAT(nuc_pert_strength, mag_pert_strength, points=[p for p in mine if p[0] != "U"],
                           unperturbed_wfn=wfn if own_u else None)
    blob = {}
    if own_u:
        E_corr, T_list = lists[-1]
        lists = lists[:-1]
        blob[("U", 0, 0)] = (E_corr, T_list)
    slot = lambda p: ((0 if p[0] == "R" else 6) + (0 if p[2] > 0 else 1), p[1])
    if world > 1:
        # exchange (C, T) of every point; basis handles are rebuilt locally (geometry only)
        for p in mine:
            if p[0] != "U":
                g, i = slot(p)
                blob[p] = (lists[g][i], lists[g + 4][i])
        lists = [list(x) for x in lists]
        rows = sorted(set(a for a, _ in owned_elements(3 * natom, rank, world)))
        for p, val in exchange_points(dist, blob, world).items():
            if p[0] == "U":
                blob[p] = val
                continue
            g, i = slot(p)
            lists[g][i], lists[g + 4][i] = val
        for p in pts[1:]:                  # basis handles for points solved elsewhere (owned rows and field points only)
            g, i = slot(p)
            if lists[g + 2][i] is None and (p[0] == "B" or i in rows):
                if p[0] == "R":
                    fd.parameters["geom"] = fd._displaced([(p[1], p[2] * nuc_pert_strength)])
                    lists[g + 2][i] = Hamiltonian(fd.parameters).basis_set
                    fd._reset()
                else:
                    lists[g + 2][i] = basis
    else:
        rows = None
    E_corr, T_list = blob[("U", 0, 0)]
    if config.VERBOSE and rank == 0:
        print("Total Energy: ", wfn.E_SCF + E_corr + wfn.H.E_nuc)
    (nuc_pos_C, nuc_neg_C, nuc_pos_basis, nuc_neg_basis, nuc_pos_T, nuc_neg_T,
     mag_pos_C, mag_neg_C, mag_pos_basis, mag_neg_basis, mag_pos_T, mag_neg_T) = lists

    AATs = AAT(parameters, wfn, C, basis, T_list, nuc_pos_C, nuc_neg_C, nuc_pos_basis, nuc_neg_basis, nuc_pos_T,
               nuc_neg_T, mag_pos_C, mag_neg_C, mag_pos_basis, mag_neg_basis, mag_pos_T, mag_neg_T,
               nuc_pert_strength, mag_pert_strength, rows=rows)
    spatial = parameters['method'] in ('RHF', 'MP2', 'CID', 'CISD')
    fn = AATs.compute_spatial_aats if spatial else AATs.compute_SO_aats
    I = np.zeros((3 * natom, 3))
    for a, b in owned_elements(3 * natom, rank, world):
        I[a, b] = fn(a, b, normalization)
    I = gather_tensor(dist, I, world)
    if config.VERBOSE and rank == 0:
        print(I, "\n")
    return I


def gather_energies(dist, values, owner, rank, world):
    """Energy-only drivers: every rank holds the energies of the points it owns; a sum-all-reduce of the
    disjointly filled float64 vector (real, imaginary parts) is the gather.  values[k] is None for points
    of other ranks."""
    n = len(owner)
    buf = np.zeros((n, 2))
    for k in range(n):
        if owner[k] == rank:
            buf[k] = (np.real(values[k]), np.imag(values[k]))
    buf = gather_tensor(dist, buf, world)
    return [complex(re, im) if im != 0.0 else float(re) for re, im in buf]


def compute_parallel_apts(parameters, nuc_pert_strength, elec_pert_strength):
    """Sharded counterpart of finite_difference.compute_APT (fin_diff.py:151-263): the 36N
    (R +- h_R) x (F +- h_F) energy points are partitioned over the ranks, every rank runs its solves
    batched on its GPU, the energies are gathered with one all-reduce and differenced on the host.
    Returns the (3N, 3) tensor on every rank."""
    dist, rank, world = _dist()
    fd = finite_difference(parameters, None, None)
    pts = fd.apt_points()
    owner = partition(pts, [1.0] * len(pts), world)
    mine = [p for p, o in zip(pts, owner) if o == rank]
    solved = dict(zip(mine, fd.solve_apt_points(mine, nuc_pert_strength, elec_pert_strength)))
    energies = gather_energies(dist, [solved.get(p) for p in pts], owner, rank, world)
    return fd.compute_APT(nuc_pert_strength, elec_pert_strength, energies=dict(zip(pts, energies)))
