"""Host-side mirror of `utilities.subdiff` (reference utilities.py:26-73) over
the CUDA soft-threshold kernel (`ecw_subdiff`).
Scope: `subdiff` is on the hot path.  Everything else here (AO-integral helpers, Koopmans guesses) serves
the harness solvers and the property targets of `exp_pot.Exp` (ecw_cc_b200/harness/__init__.py), not the product path.
"""
import numpy as np

from ._lib import lib, EcwError


def subdiff(eq, var, alpha, R_format=False):
    """Sub-gradient of the L1-regularised functional, element-wise:
    `var > 0 -> eq + alpha`, otherwise soft-threshold of `eq` (the reference's
    second loop runs over `var <= 0`, utilities.py:59-67)."""
    import torch
    if R_format:
        raise NotImplementedError("R_format conversion is broken in the reference as well (utilities.py:27)")
    is_t = isinstance(eq, torch.Tensor)
    shape = tuple(eq.shape)
    if shape != tuple(var.shape):  # utilities.py:39-40
        raise ValueError('equations and variables matrices must have the same shape')
    if not torch.cuda.is_available():
        raise EcwError("no CUDA device: subdiff has no CPU implementation in ecw_cc_b200")
    dev = eq.device if is_t else torch.device("cuda", torch.cuda.current_device())
    e = eq if is_t else torch.from_numpy(np.ascontiguousarray(eq, dtype=np.float64)).to(dev)
    v = var if isinstance(var, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(var, dtype=np.float64)).to(dev)
    e = e.contiguous()
    v = v.contiguous()
    out = torch.empty_like(e)
    st = torch.cuda.current_stream(dev).cuda_stream
    if lib.ecw_subdiff(e.data_ptr(), v.data_ptr(), float(alpha), out.data_ptr(), e.numel(), st) != 0:
        raise EcwError("ecw_subdiff failed")
    return out if is_t else out.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------------
# Host-side helpers around the path (n x n and o x v work; numpy).  Restated from the reference's `utilities` for the
# callers mirrored in this package (`exp_pot.Exp`, `Solver_ES`): same names, arguments and conventions.
# ---------------------------------------------------------------------------------------------------------------------

def convert_r_to_g_rdm1(rdm_r):
    """AO matrix in spatial (R) format -> spin-blocked G format, each spin block carrying HALF of it
    (utilities.py:226-243)."""
    nao = rdm_r.shape[0]
    g = np.zeros((2 * nao, 2 * nao))
    g[:nao, :nao] = g[nao:, nao:] = 0.5 * np.asarray(rdm_r)
    return g


def convert_g_to_ru_rdm1(rdm1_g):
    """Spin-blocked AO rdm1 -> (alpha + beta, (alpha, beta)) (utilities.py:189-206)."""
    nao = rdm1_g.shape[0] // 2
    a, b = rdm1_g[:nao, :nao], rdm1_g[nao:, nao:]
    return a + b, (a, b)


def ao_to_mo(rdm1_ao, mo_coeff):
    """C^-1 X C^-T (utilities.py:361-378)."""
    if rdm1_ao.shape != mo_coeff.shape:
        raise ValueError('Rdm1 and MOs coefficients must have the same dimension')
    inv = np.linalg.inv(mo_coeff)
    return inv @ rdm1_ao @ inv.T


def mo_to_ao(rdm1_mo, mo_coeff):
    """C X C^T (utilities.py:381-394)."""
    if rdm1_mo.shape != mo_coeff.shape:
        raise ValueError('rdm1 and mo coeff must have the same size')
    return mo_coeff @ rdm1_mo @ mo_coeff.T


def convert_aoint(int_ao, mo_coeff, g=True):
    """AO one-electron integrals -> the spin-orbital matrices A_pq that dress the Fock matrix (utilities.py:311-340).
    The reference pushes the integrals through its rdm1 transformation (half per spin block, C^-1 A C^-T); the fit's
    weights L absorb that convention, so it is kept as is.  A leading dimension of 3 means x, y, z components."""
    if not g:
        nao = mo_coeff.shape[0]
        mo = np.zeros((2 * nao, 2 * nao))
        mo[:nao, 0::2] = mo_coeff
        mo[nao:, 1::2] = mo_coeff
    else:
        mo = mo_coeff
    int_ao = np.asarray(int_ao)
    if int_ao.shape[0] == 3:
        return np.stack([ao_to_mo(convert_r_to_g_rdm1(x), mo) for x in int_ao])
    return ao_to_mo(convert_r_to_g_rdm1(int_ao), mo)


def _expectation(mol, rdm1, g, aobasis, mo_coeff, ints, name, comp=None):
    if aobasis is False:
        if mo_coeff is None:
            raise ValueError('mo_coeff must be given if rdm is not in AOs basis')
        rdm1 = mo_coeff @ rdm1 @ mo_coeff.T
    if g:
        rdm1 = convert_g_to_ru_rdm1(rdm1)[0]
    if ints is None:
        if name == 'int1e_r':
            charges, coords = mol.atom_charges(), mol.atom_coords()
            with mol.with_common_orig(np.einsum('z,zr->r', charges, coords) / charges.sum()):
                ints = mol.intor_symmetric(name, comp=3)
        else:
            ints = mol.intor_symmetric(name)
    return np.einsum('xij,ji->x', ints, rdm1) if comp else np.einsum('ij,ji', ints, rdm1)


def Ekin(mol, rdm1, g=True, aobasis=True, mo_coeff=None, ek_int=None):
    """tr(T rdm1) (utilities.py:985-1014)."""
    return _expectation(mol, rdm1, g, aobasis, mo_coeff, ek_int, 'int1e_kin')


def v1e(mol, rdm1, g=True, aobasis=True, mo_coeff=None, v1e_int=None):
    """tr(V_ne rdm1) (utilities.py:1017-1046)."""
    return _expectation(mol, rdm1, g, aobasis, mo_coeff, v1e_int, 'int1e_nuc')


def dipole(mol, rdm1, g=True, aobasis=True, mo_coeff=None, dip_int=None):
    """Electronic dipole components about the centre of nuclear charge (utilities.py:1049-1086)."""
    return _expectation(mol, rdm1, g, aobasis, mo_coeff, dip_int, 'int1e_r', comp=3)


def convert_r_to_g_amp(amp):
    """o x v singles in spatial format -> spin-orbital [a b a b ..] format, same-spin blocks (utilities.py:137-157)."""
    amp = np.asarray(amp)
    if amp.ndim != 2:
        raise ValueError('only singles are converted without PySCF')
    return np.kron(amp, np.eye(2))


def koopman_init_guess(mo_energy, mo_occ, nstates=[1, 0], koop_idx=None, core_ene_thresh=10.):
    """Koopmans start vectors r1 (spin-orbital format) and their orbital-energy differences (utilities.py:397-478):
    the k-th lowest valence (core) single excitation, shifted by koop_idx, as ONE spin-orbital element — the
    same-spin pair of the spatial excitation with its first element removed (valence), or, for core states, with the
    two ROWS indexed by the first element's (row, column) removed, which is what the reference's indexing does."""
    if koop_idx is not None and sum(nstates) != len(koop_idx):
        raise ValueError('Number of given Koopman indices should be equal to the number of excited states')
    nval, ncor = nstates
    val_idx = (np.zeros(nval, dtype=int) if koop_idx is None else koop_idx[:nval]) if nval != 0 else [0]
    core_idx = (np.zeros(ncor, dtype=int) if koop_idx is None else koop_idx[nval:]) if ncor != 0 else [0]
    mo_energy, mo_occ = np.asarray(mo_energy)[0::2], np.asarray(mo_occ)[0::2]
    occidx, viridx = np.where(mo_occ == 1)[0], np.where(mo_occ == 0)[0]
    nocc, nvir = occidx.shape[0], viridx.shape[0]
    ncore = np.where(abs(mo_energy[:nocc]) > core_ene_thresh)[0].shape[0]
    e_ia = mo_energy[viridx] - mo_energy[occidx, None]
    eia_val, eia_core = e_ia[ncore:, :].ravel(), e_ia[:ncore, :].ravel()
    if nval > eia_val.size or ncor > eia_core.size:
        raise Warning('The size of the basis is smaller than the number of requested states')
    x0, DE = [], []
    order = np.argsort(eia_val)
    for i in range(min(nval, eia_val.size)):
        k = order[i + val_idx[i]]
        r = np.zeros(eia_val.size)
        r[k] = 1
        r = convert_r_to_g_amp(np.vstack((np.zeros((ncore, nvir)), r.reshape(nocc - ncore, nvir))))
        r[tuple(np.transpose(np.nonzero(r))[0])] = 0
        x0.append(r)
        DE.append(eia_val[k])
    order = np.argsort(eia_core)
    for i in range(min(ncor, eia_core.size)):
        k = order[i + core_idx[i]]
        r = np.zeros(eia_core.size)
        r[k] = 1
        r = convert_r_to_g_amp(np.vstack((r.reshape(ncore, nvir), np.zeros((nocc - ncore, nvir)))))
        r[np.transpose(np.nonzero(r))[0]] = 0
        x0.append(r)
        DE.append(eia_core[k])
    return x0, DE


def get_DE(mo_energy, rs):
    """Orbital-energy difference of the largest element of rs (utilities.py:481-493)."""
    nocc, nvir = rs.shape
    eia = mo_energy[nocc:] - mo_energy[:nocc, None]
    return eia[np.unravel_index(np.argmax(rs), (nocc, nvir))]


def check_spin(amp_r, amp_l):
    """sum_ia r_ia l_ia s_ia with s = -1 (alpha->beta), +1 (beta->alpha), 0 otherwise (utilities.py:551-571)."""
    s = np.zeros_like(amp_r)
    s[::2, 1::2] = -1
    s[1::2, 0::2] = 1
    return np.einsum('ia,ia,ia', amp_r, amp_l, s)


def get_norm(rs, ls, r0, l0):
    """l0 r0 + <r|l> (utilities.py:625-642)."""
    if rs.shape != ls.shape:
        raise ValueError('Shape of both set of amplitudes must be the same')
    return l0 * np.conjugate(r0) + np.sum(np.conjugate(rs) * ls)


def check_ortho(rn, ln, r0n, l0n):
    """Matrix of (<k|l> + <l|k>)/2 for lists of states (utilities.py:730-758)."""
    ns = len(rn)
    if ns != len(ln):
        raise ValueError('r and l list of vectors must be the same length')
    C = np.zeros((ns, ns))
    for k in range(ns):
        for l in range(ns):
            C[k, l] = np.ravel((get_norm(rn[k], ln[l], r0n[k], l0n[l]) + get_norm(rn[l], ln[k], r0n[l], l0n[k])) / 2.)[0]
    return C
