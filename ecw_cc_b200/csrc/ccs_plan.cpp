// ccs_plan.cpp — plans of the ECW-CCS intermediates (reference: CCS.py:271-312 T1inter, :490-537 L1inter, :774-872
// R1inter, :1164-1234 es_L1inter): everything of a CCS ground- or excited-state iteration that touches an integral
// block.  One cached plan per function (and presence of the state potential), i.e. one C call — and on one GPU one CUDA
// graph launch — instead of 10-25 primitive calls.  The singles-sized updates that consume them stay sequences of
// the primitive ops (host-side scalars between them: Em extraction, r0 / l0).
//
// Slots: ts = "t1", fsp = "fsp", state potential vm = "fock" (n x n, only read when has_vm), results:
//   F  = "rdm1" slot, an n x n matrix whose blocks carry the one-body intermediates (which block holds what is stated
//        per builder; blocks that a function does not define are zero),
//   W  = "out2" slot viewed [v,o,o,v]  (Wbija / Wakic),
//   X  = "out1" slot [o,v]  (second singles-sized result: Pia / P),
//   scal[0] = the energy-like scalar (E / Er / El).
// Integral blocks other than the canonical ones follow Eris.py:128:  ovvo[jabi] = -ovov_ph[iajb],
// voov[bija] = -ovov_ph[jbia], oovo[kjbi] = -ooov[kjib], vovv[bica] = -ovvv[ibca], ovov_ph[(ia),(nf)] = ovov[naif].
#include "ccsd_plan_detail.h"

namespace ecw {

using namespace detail;

namespace {

struct CcsSlots : Slots {
  Tensor ts, vm, F, Foo, Fov, Fvo, Fvv, W, X, voo, vov, vvo, vvv;
  explicit CcsSlots(const Sizes& z) : Slots(z) {
    ts = t1;
    vm = fock;
    F = rdm1;
    Foo = block2(F, 0, o, 0, o);
    Fov = block2(F, 0, o, o, v);
    Fvo = block2(F, o, v, 0, o);
    Fvv = block2(F, o, v, o, v);
    W = make_tensor(S_OUT2, 0, {v, o, o, v});
    X = out1;
    voo = block2(vm, 0, o, 0, o);
    vov = block2(vm, 0, o, o, v);
    vvo = block2(vm, o, v, 0, o);
    vvv = block2(vm, o, v, o, v);
  }
};

// scal[k] = cf <ts, fov> + cg <ts, G>,  G[jb] = sum_kc ts[kc] oovv[jkbc]   (CCS.py:226-249 and the E terms of the inters)
void emit_ccs_energy(Plan& P, const CcsSlots& s, double cf, double cg, int k) {
  Tensor f = P.tmp({s.o, s.v});
  P.axpby(1.0, s.fov, 0.0, f);
  P.dot(cf, f, s.ts, 0.0, k);
  P.contract(1.0, s.ts, "kc", s.oovv, "jkbc", 0.0, f, "jb", "G");
  P.dot(cg, s.ts, f, 1.0, k);
  P.release(f);
}

}  // namespace

// F: vv = Fab, oo = Fji, vo = Fai.
void build_ccs_t1inter(Plan& P, const Sizes& z) {
  z.apply(P);
  CcsSlots s(z);
  const int64_t o = s.o, v = s.v;
  P.fill(s.F, 0.0);
  Tensor Fai = P.tmp({v, o});
  P.axpby(1.0, s.fvo, 0.0, Fai);
  P.contract(-1.0, s.ts, "jb", s.ovov_ph, "iajb", 1.0, Fai, "ai", "Fai ovvo");          // 'jb,jabi->ai' ovvo
  P.axpby(1.0, Fai, 0.0, s.Fvo);
  P.release(Fai);
  Tensor Fab = P.tmp({v, v});
  P.axpby(1.0, s.fvv, 0.0, Fab);
  P.contract(-1.0, s.fov, "jb", s.ts, "ja", 1.0, Fab, "ab");
  P.contract(1.0, s.ts, "jc", s.ovvv, "jacb", 1.0, Fab, "ab", "Fab ovvv");
  P.axpby(1.0, Fab, 0.0, s.Fvv);
  P.release(Fab);
  Tensor Fji = P.tmp({o, o});
  P.axpby(1.0, s.foo, 0.0, Fji);
  P.contract(-1.0, s.ts, "kb", s.ooov, "kjib", 1.0, Fji, "ji", "Fji oovo");             // 'kb,kjbi->ji' oovo
  Tensor x = P.tmp({o, v});
  P.contract(1.0, s.ts, "kc", s.oovv, "jkcb", 0.0, x, "jb");
  P.contract(-1.0, s.ts, "ib", x, "jb", 1.0, Fji, "ji");
  P.release(x);
  P.axpby(1.0, Fji, 0.0, s.Foo);
  P.release(Fji);
}

// F: ov = Fia, vv = Fba, oo = Fij;  W = Wbija;  scal[0] = E (only with the E term).
void build_ccs_l1inter(Plan& P, const Sizes& z, int e_term) {
  z.apply(P);
  CcsSlots s(z);
  const int64_t o = s.o, v = s.v;
  P.fill(s.F, 0.0);
  Tensor Fba = P.tmp({v, v});
  P.axpby(1.0, s.fvv, 0.0, Fba);
  P.contract(-1.0, s.fov, "ja", s.ts, "jb", 1.0, Fba, "ba");
  P.contract(1.0, s.ovvv, "jbca", s.ts, "jc", 1.0, Fba, "ba", "Fba ovvv");
  Tensor x = P.tmp({o, v});
  P.contract(1.0, s.oovv, "jkca", s.ts, "jc", 0.0, x, "ka");
  P.contract(-1.0, x, "ka", s.ts, "kb", 1.0, Fba, "ba");
  P.axpby(1.0, Fba, 0.0, s.Fvv);
  P.release(Fba);
  Tensor Fij = P.tmp({o, o});
  P.axpby(1.0, s.foo, 0.0, Fij);
  P.contract(1.0, s.fov, "ib", s.ts, "jb", 1.0, Fij, "ij");
  P.contract(-1.0, s.ooov, "kijb", s.ts, "kb", 1.0, Fij, "ij", "Fij oovo");             // 'kibj,kb->ij' oovo
  P.contract(1.0, s.oovv, "kibc", s.ts, "kb", 0.0, x, "ic");
  P.contract(1.0, x, "ic", s.ts, "jc", 1.0, Fij, "ij");
  P.axpby(1.0, Fij, 0.0, s.Foo);
  P.release(Fij);
  P.permute(-1.0, s.ovov_ph, "jbia", 0.0, s.W, "bija", "W voov");                       // voov[bija] = -ovov_ph[jbia]
  P.contract(-1.0, s.ooov, "kija", s.ts, "kb", 1.0, s.W, "bija", "W ooov");
  Tensor y = P.tmp({o, v, v, v});
  P.contract(1.0, s.oovv, "kica", s.ts, "kb", 0.0, y, "icab", "W oovv.ts");
  P.contract(-1.0, y, "icab", s.ts, "jc", 1.0, s.W, "bija", "W oovv.ts.ts");
  P.release(y);
  P.contract(-1.0, s.ovvv, "ibca", s.ts, "jc", 1.0, s.W, "bija", "W vovv");             // 'bica,jc->bija' vovv
  P.axpby(1.0, s.fov, 0.0, x);
  P.contract(1.0, s.oovv, "jiba", s.ts, "jb", 1.0, x, "ia");
  P.axpby(1.0, x, 0.0, s.Fov);
  P.release(x);
  if (e_term) emit_ccs_energy(P, s, -1.0, -0.5, 0);       // without it the scalar is not written (the caller reports 0)
}

// F: vv = Fab, oo = Fji, ov = Tia;  W = Wakic;  X = Pia;  scal[0] = Er.
void build_ccs_r1inter(Plan& P, const Sizes& z, int has_vm) {
  z.apply(P);
  CcsSlots s(z);
  const int64_t o = s.o, v = s.v;
  P.fill(s.F, 0.0);
  Tensor x = P.tmp({o, v});
  Tensor Fab = P.tmp({v, v});
  P.axpby(1.0, s.fvv, 0.0, Fab);
  P.contract(-1.0, s.ts, "ja", s.fov, "jb", 1.0, Fab, "ab");
  P.contract(1.0, s.ts, "jc", s.ovvv, "jacb", 1.0, Fab, "ab", "Fab ovvv");
  P.contract(1.0, s.ts, "jc", s.oovv, "jkcb", 0.0, x, "kb");                            // 'jc,ka,jkcb->ab'
  P.contract(-1.0, s.ts, "ka", x, "kb", 1.0, Fab, "ab");
  P.axpby(1.0, Fab, 0.0, s.Fvv);
  P.release(Fab);
  Tensor Fji = P.tmp({o, o});
  P.axpby(1.0, s.foo, 0.0, Fji);
  P.contract(1.0, s.ts, "ib", s.fov, "jb", 1.0, Fji, "ji");
  P.contract(-1.0, s.ts, "kb", s.ooov, "kjib", 1.0, Fji, "ji", "Fji oovo");
  P.contract(1.0, s.ts, "kb", s.oovv, "kjbc", 0.0, x, "jc");                            // 'kb,ic,kjbc->ji'
  P.contract(1.0, s.ts, "ic", x, "jc", 1.0, Fji, "ji");
  P.axpby(1.0, Fji, 0.0, s.Foo);
  P.release(Fji);
  P.permute(-1.0, s.ovov_ph, "iakc", 0.0, s.W, "akic", "W voov");                       // voov[akic]
  P.contract(-1.0, s.ts, "ib", s.ovvv, "kabc", 1.0, s.W, "akic", "W vovv");             // vovv[akbc]
  Tensor zt = P.tmp({o, o, o, v});
  P.contract(1.0, s.ts, "ib", s.oovv, "jkbc", 0.0, zt, "ijkc", "W ts.oovv");            // 'ib,ja,jkbc->akic'
  P.contract(-1.0, s.ts, "ja", zt, "ijkc", 1.0, s.W, "akic", "W ts.ts.oovv");
  P.contract(-1.0, s.ts, "ja", s.ooov, "jkic", 1.0, s.W, "akic", "W ooov");
  emit_ccs_energy(P, s, 1.0, 0.5, 0);
  // Tia = Zai^T + ts.Zab - Zji.ts
  Tensor Zab = P.tmp({v, v});
  P.axpby(1.0, s.fvv, 0.0, Zab);
  P.contract(-1.0, s.ts, "ja", s.fov, "jb", 1.0, Zab, "ab");
  Tensor Zji = P.tmp({o, o});
  P.axpby(1.0, s.foo, 0.0, Zji);
  P.contract(-1.0, s.ts, "kb", s.ooov, "kjib", 1.0, Zji, "ji");
  P.contract(1.0, s.ts, "ic", s.oovv, "jkbc", 0.0, zt, "ijkb");
  P.contract(-1.0, s.ts, "kb", zt, "ijkb", 1.0, Zji, "ji");
  P.release(zt);
  Tensor Zai = P.tmp({v, o});
  P.axpby(1.0, s.fvo, 0.0, Zai);
  P.contract(-1.0, s.ts, "jb", s.ovov_ph, "iajb", 1.0, Zai, "ai", "Zai ovvo");
  Tensor u = P.tmp({o, v, v, o});
  P.contract(1.0, s.ovvv, "jabc", s.ts, "ic", 0.0, u, "jabi", "Zai ovvv.ts");           // 'jb,ic,jabc->ai'
  P.contract(1.0, s.ts, "jb", u, "jabi", 1.0, Zai, "ai");
  P.release(u);
  P.permute(1.0, Zai, "ai", 0.0, x, "ia");
  P.contract(1.0, s.ts, "ib", Zab, "ab", 1.0, x, "ia");
  P.contract(-1.0, s.ts, "ja", Zji, "ji", 1.0, x, "ia");
  P.axpby(1.0, x, 0.0, s.Fov);
  P.release(Zai);
  P.release(Zji);
  P.release(Zab);
  // Pia (CCS.py:861-872): -(v_vo + v_vv.ts) and the literal 'ii,ja,ib->ai' term (Q9), all with vm -> -vm
  if (!has_vm) {
    P.fill(s.X, 0.0);
  } else {
    Tensor Pai = P.tmp({v, o});
    P.axpby(-1.0, s.vvo, 0.0, Pai);
    P.contract(-1.0, s.vvv, "ab", s.ts, "ib", 1.0, Pai, "ai");
    Tensor ones = P.tmp({std::max(o, v)});
    P.fill(ones, 1.0);
    Tensor ones_o = ones, ones_v = ones;
    ones_o.dim[0] = o;
    ones_v.dim[0] = v;
    Tensor col = P.tmp({v}), row = P.tmp({o});
    P.contract(1.0, s.ts, "ja", ones_o, "j", 0.0, col, "a");
    P.contract(1.0, s.ts, "ib", ones_v, "b", 0.0, row, "i");
    // drow[i] = (-vm)[i,i] row[i]: the diagonal of the outer product of the two vectors
    Tensor diag = make_tensor(S_FOCK, 0, {o});
    diag.str[0] = s.n + 1;
    Tensor outer = P.tmp({o, o});
    P.contract(-1.0, diag, "i", row, "k", 0.0, outer, "ik");
    Tensor drow = outer;
    drow.nd = 1;
    drow.dim[0] = o;
    drow.str[0] = o + 1;
    P.contract(-1.0, col, "a", drow, "i", 1.0, Pai, "ai");
    P.permute(1.0, Pai, "ai", 0.0, s.X, "ia");
    P.release(outer);
    P.release(row);
    P.release(col);
    P.release(ones);
    P.release(Pai);
  }
  P.release(x);
}

// F: vv = Fba, oo = Fij, ov = Zia;  W = Wbija;  X = P;  scal[0] = El.
void build_ccs_esl1inter(Plan& P, const Sizes& z, int has_vm) {
  z.apply(P);
  CcsSlots s(z);
  const int64_t o = s.o, v = s.v;
  P.fill(s.F, 0.0);
  Tensor x = P.tmp({o, v});
  Tensor Fba = P.tmp({v, v});
  P.axpby(1.0, s.fvv, 0.0, Fba);
  P.contract(-1.0, s.ts, "jb", s.fov, "ja", 1.0, Fba, "ba");
  P.contract(1.0, s.ts, "jc", s.ovvv, "jbca", 1.0, Fba, "ba", "Fba ovvv");
  P.contract(1.0, s.ts, "jc", s.oovv, "jkca", 0.0, x, "ka");                            // 'jc,kb,jkca->ba'
  P.contract(-1.0, s.ts, "kb", x, "ka", 1.0, Fba, "ba");
  P.axpby(1.0, Fba, 0.0, s.Fvv);
  P.release(Fba);
  Tensor Fij = P.tmp({o, o});
  P.axpby(1.0, s.foo, 0.0, Fij);
  P.contract(1.0, s.ts, "jb", s.fov, "ib", 1.0, Fij, "ij");
  P.contract(-1.0, s.ts, "kb", s.ooov, "kijb", 1.0, Fij, "ij", "Fij oovo");             // oovo[kibj]
  P.contract(1.0, s.ts, "kb", s.oovv, "kibc", 0.0, x, "ic");                            // 'kb,jc,kibc->ij'
  P.contract(1.0, s.ts, "jc", x, "ic", 1.0, Fij, "ij");
  P.axpby(1.0, Fij, 0.0, s.Foo);
  P.release(Fij);
  P.permute(-1.0, s.ovov_ph, "jbia", 0.0, s.W, "bija", "W voov");
  P.contract(-1.0, s.ts, "kb", s.ooov, "kija", 1.0, s.W, "bija", "W ooov");
  P.contract(-1.0, s.ts, "jc", s.ovvv, "ibca", 1.0, s.W, "bija", "W vovv");             // vovv[bica]
  Tensor y = P.tmp({v, o, v, v});
  P.contract(1.0, s.ts, "kb", s.oovv, "kica", 0.0, y, "bica", "W ts.oovv");             // 'jc,kb,kica->bija'
  P.contract(-1.0, s.ts, "jc", y, "bica", 1.0, s.W, "bija", "W ts.ts.oovv");
  P.release(y);
  emit_ccs_energy(P, s, 1.0, 0.5, 0);
  P.axpby(1.0, s.fov, 0.0, x);
  P.contract(1.0, s.ts, "jb", s.oovv, "jiba", 1.0, x, "ia");
  P.axpby(1.0, x, 0.0, s.Fov);
  P.release(x);
  if (has_vm) P.axpby(-1.0, s.vov, 0.0, s.X);
  else P.fill(s.X, 0.0);
}

}  // namespace ecw
