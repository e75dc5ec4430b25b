// raco_lsodes.cpp -- ORACLE (test infrastructure only, see raco.h): restatement of
// ODEPACK's DLSODES as the reference calls it (MF = 21: BDF, chord iteration with a
// user-supplied sparse Jacobian, IA/JA given; ITOL = 4; ITASK = 1 or 4; IOPT = 1).
//   driver      <- src/opkdmain.f:3069-3588
//   dstode()    <- DSTODE  src/opkda1.f:629-1126   (label numbers kept in comments)
//   dprjs()     <- DPRJS   src/opkda1.f:1664-1863  (MITER = 1 branch)
//   dsolss()    <- DSOLSS  src/opkda1.f:1864-1942
//   dcfode()    <- DCFODE  src/opkda1.f:46-173     (METH = 2)
//   dintdy()    <- DINTDY  src/opkda1.f:174-281    (K = 0)
//   dewset/dvnorm <- src/opkda1.f:1127-1209
// Sparse LU: raco_sparse.cpp (YSMP's role).  The matrix P is kept and, when the
// Jacobian is reused, rescaled in place exactly as DPRJS does (label 250).
#include "raco_internal.hpp"
#include <algorithm>
#include <limits>

namespace raco {

#define YHc(i, j) YH[(size_t)((j) - 1) * NYH + ((i) - 1)]  // YH(i,j), 1-based

void Lsodes::dcfode() {  // METH = 2 branch, labels 200-230
  double PC[13];
  PC[1] = 1.0;
  double RQ1FAC = 1.0;
  for (int nq = 1; nq <= 5; ++nq) {
    double FNQ = nq;
    int NQP1 = nq + 1;
    PC[NQP1] = 0.0;
    for (int IB = 1; IB <= nq; ++IB) {
      int I = nq + 2 - IB;
      PC[I] = PC[I - 1] + FNQ * PC[I];
    }
    PC[1] = FNQ * PC[1];
    for (int I = 1; I <= NQP1; ++I) ELCO[I][nq] = PC[I] / PC[2];
    ELCO[2][nq] = 1.0;
    TESCO[1][nq] = RQ1FAC;
    TESCO[2][nq] = NQP1 / ELCO[1][nq];
    TESCO[3][nq] = (nq + 2) / ELCO[1][nq];
    RQ1FAC = RQ1FAC / FNQ;
  }
}

double Lsodes::dvnorm(const double* v, const double* w) const {
  double sum = 0.0;
  for (int i = 0; i < N; ++i) { double a = v[i] * w[i]; sum += a * a; }
  return std::sqrt(sum / N);
}

void Lsodes::dintdy(double t, double* dky) const {  // K = 0
  double S = (t - TN) / H;
  for (int i = 1; i <= N; ++i) dky[i - 1] = YHc(i, L);
  for (int JB = 1; JB <= NQ; ++JB) {
    int J = NQ - JB, JP1 = J + 1;
    for (int i = 1; i <= N; ++i) dky[i - 1] = YHc(i, JP1) + S * dky[i - 1];
  }
}

void Lsodes::dsolss(double* x) {
  IERSL = 0;
  lu.solve(x);
  ++n_solve;
}

void Lsodes::dprjs(double* y) {
  const int nnz_user = (int)ja.size();
  double HL0 = H * EL0;
  double CON = -HL0;
  int JOK = 1;
  if (NST == 0 || NST >= NSLJ + MSBJ) JOK = 0;
  if (ICF == 1 && std::fabs(RC - 1.0) < CCMXJ) JOK = 0;
  if (ICF == 2) JOK = 0;
  bool reeval = (JOK == 0);
  if (!reeval) {
    // label 250: reconstruct new P from old P
    JCUR = 0;
    double RCON = CON / CON0;
    double RCONT = std::fabs(CON) / CONMIN;
    if (RCONT > RBIG && iplost) reeval = true;
    else {
      for (int j = 1; j <= N; ++j) {
        for (int k = ia[j - 1] - 1; k < ia[j] - 1; ++k) {
          int i = ja[k];
          double PIJ = Pval[k];
          if (i == j) {
            PIJ = PIJ - 1.0;
            if (std::fabs(PIJ) < PSMALL) { iplost = 1; CONMIN = std::min(std::fabs(CON0), CONMIN); }
          }
          PIJ = PIJ * RCON;
          if (i == j) PIJ = PIJ + 1.0;
          Pval[k] = PIJ;
        }
      }
      // diagonals appended by DPREP (absent from the user's pattern) hold P_jj too
      {
        double PIJ = Padd - 1.0;
        if (lu.nnz_a > nnz_user && std::fabs(PIJ) < PSMALL) { iplost = 1; CONMIN = std::min(std::fabs(CON0), CONMIN); }
        Padd = PIJ * RCON + 1.0;
      }
    }
  }
  if (reeval) {
    // label 20/100
    JCUR = 1;
    NJE = NJE + 1;
    NSLJ = NST;
    iplost = 0;
    CONMIN = std::fabs(CON);
    if (jac_csc) {
      jac_csc(TN, y, Jval.data());
      for (int j = 1; j <= N; ++j)
        for (int k = ia[j - 1] - 1; k < ia[j] - 1; ++k) {
          Pval[k] = Jval[k] * CON;
          if (ja[k] == j) Pval[k] += 1.0;
        }
    } else {
      for (int j = 1; j <= N; ++j) {
        for (int i = 0; i < N; ++i) FTEM[i] = 0.0;
        jac_col(TN, y, j, FTEM.data());
        for (int k = ia[j - 1] - 1; k < ia[j] - 1; ++k) {
          int i = ja[k];
          Pval[k] = FTEM[i - 1] * CON;
          if (i == j) Pval[k] += 1.0;
        }
      }
    }
    Padd = 1.0;  // J entry of an appended diagonal is FTEM(j) = 0 unless JAC fills it; see note
    if (!jac_csc && lu.nnz_a > nnz_user) {
      // DPRJS reads FTEM(I) for every entry of the EXTENDED pattern, including the
      // appended diagonals: P_jj = FTEM(j)*CON + 1.  For the chemistry network those
      // species are never consumed, so FTEM(j) = 0 there; the generic path keeps 1.
    }
  }
  // label 290
  NLU = NLU + 1;
  CON0 = CON;
  IERPJ = 0;
  int flag = lu.factor(Pval.data(), Padd);
  if (flag != 0) IERPJ = 1;
}

void Lsodes::dewset(const double* rtol, const double* atol) {
  for (int i = 0; i < N; ++i) EWT[i] = rtol[i] * std::fabs(YH[i]) + atol[i];
}

bool Lsodes::ewt_invert_ok() {
  for (int i = 0; i < N; ++i) {
    if (EWT[i] <= 0.0) { IMXER = i + 1; return false; }
    EWT[i] = 1.0 / EWT[i];
  }
  return true;
}

void Lsodes::dstode(double* Y) {
  double* YH1 = YH.data() - 1;  // YH1(I), 1-based flat
  double DCON, DDN, DEL = 0, DELP, DSM = 0, DUP, EXDN, EXSM, EXUP, R, RH = 0, RHDN, RHSM, RHUP, TOLD;
  int I1, IREDO = 0, IRET = 0, M = 0, NCF, NEWQ = 0;
  KFLAG = 0;
  TOLD = TN;
  NCF = 0;
  IERPJ = 0;
  IERSL = 0;
  JCUR = 0;
  ICF = 0;
  DELP = 0.0;
  if (JSTART > 0) goto L200;
  if (JSTART == -1) goto L100;
  if (JSTART == -2) goto L160;
  LMAX = MAXORD + 1;
  NQ = 1;
  L = 2;
  IALTH = 2;
  RMAX = 10000.0;
  RC = 0.0;
  EL0 = 1.0;
  CRATE = 0.7;
  HOLD = H;
  MEO = METH;
  NSLP = 0;
  IPUP = MITER;
  IRET = 3;
  goto L140;
L100:
  IPUP = MITER;
  LMAX = MAXORD + 1;
  if (IALTH == 1) IALTH = 2;
  // METH == MEO always (BDF only)
  if (NQ <= MAXORD) goto L160;
  // MAXORD is never reduced below NQ by the reference's caller
  goto L160;
L140:
  dcfode();
L150:
  for (int i = 1; i <= L; ++i) EL[i] = ELCO[i][NQ];
  NQNYH = NQ * NYH;
  RC = RC * EL[1] / EL0;
  EL0 = EL[1];
  CONIT = 0.5 / (NQ + 2);
  switch (IRET) { case 1: goto L160; case 2: goto L170; default: goto L200; }
L160:
  if (H == HOLD) goto L200;
  RH = H / HOLD;
  H = HOLD;
  IREDO = 3;
  goto L175;
L170:
  RH = std::max(RH, HMIN / std::fabs(H));
L175:
  RH = std::min(RH, RMAX);
  RH = RH / std::max(1.0, std::fabs(H) * HMXI * RH);
  R = 1.0;
  for (int j = 2; j <= L; ++j) {
    R = R * RH;
    for (int i = 1; i <= N; ++i) YHc(i, j) = YHc(i, j) * R;
  }
  H = H * RH;
  RC = RC * RH;
  IALTH = L;
  if (IREDO == 0) goto L690;
L200:
  if (std::fabs(RC - 1.0) > CCMAX) IPUP = MITER;
  if (NST >= NSLP + MSBP) IPUP = MITER;
  TN = TN + H;
  I1 = NQNYH + 1;
  for (int JB = 1; JB <= NQ; ++JB) {
    I1 = I1 - NYH;
    for (int i = I1; i <= NQNYH; ++i) YH1[i] = YH1[i] + YH1[i + NYH];
  }
L220:
  M = 0;
  for (int i = 1; i <= N; ++i) Y[i - 1] = YHc(i, 1);
  f(TN, Y, SAVF.data());
  NFE = NFE + 1;
  if (IPUP <= 0) goto L250;
  dprjs(Y);
  IPUP = 0;
  RC = 1.0;
  NSLP = NST;
  CRATE = 0.7;
  if (IERPJ != 0) goto L430;
L250:
  for (int i = 0; i < N; ++i) ACOR[i] = 0.0;
L270:
  // chord method, label 350
  for (int i = 1; i <= N; ++i) Y[i - 1] = H * SAVF[i - 1] - (YHc(i, 2) + ACOR[i - 1]);
  dsolss(Y);
  if (IERSL < 0) goto L430;
  if (IERSL > 0) goto L410;
  DEL = dvnorm(Y, EWT.data());
  for (int i = 1; i <= N; ++i) {
    ACOR[i - 1] = ACOR[i - 1] + Y[i - 1];
    Y[i - 1] = YHc(i, 1) + EL[1] * ACOR[i - 1];
  }
  // label 400
  if (M != 0) CRATE = std::max(0.2 * CRATE, DEL / DELP);
  DCON = DEL * std::min(1.0, 1.5 * CRATE) / (TESCO[2][NQ] * CONIT);
  if (DCON <= 1.0) goto L450;
  M = M + 1;
  if (M == MAXCOR) goto L410;
  if (M >= 2 && DEL > 2.0 * DELP) goto L410;
  DELP = DEL;
  f(TN, Y, SAVF.data());
  NFE = NFE + 1;
  goto L270;
L410:
  if (JCUR == 1) goto L430;
  ICF = 1;
  IPUP = MITER;
  goto L220;
L430:
  ICF = 2;
  NCF = NCF + 1;
  ++n_cfail;
  RMAX = 2.0;
  TN = TOLD;
  I1 = NQNYH + 1;
  for (int JB = 1; JB <= NQ; ++JB) {
    I1 = I1 - NYH;
    for (int i = I1; i <= NQNYH; ++i) YH1[i] = YH1[i] - YH1[i + NYH];
  }
  if (IERPJ < 0 || IERSL < 0) goto L680;
  if (std::fabs(H) <= HMIN * 1.00001) goto L670;
  if (NCF == MXNCF) goto L670;
  RH = 0.25;
  IPUP = MITER;
  IREDO = 1;
  goto L170;
L450:
  JCUR = 0;
  if (M == 0) DSM = DEL / TESCO[2][NQ];
  if (M > 0) DSM = dvnorm(ACOR.data(), EWT.data()) / TESCO[2][NQ];
  if (DSM > 1.0) goto L500;
  KFLAG = 0;
  IREDO = 0;
  NST = NST + 1;
  HU = H;
  NQU = NQ;
  for (int j = 1; j <= L; ++j)
    for (int i = 1; i <= N; ++i) YHc(i, j) = YHc(i, j) + EL[j] * ACOR[i - 1];
  IALTH = IALTH - 1;
  if (IALTH == 0) goto L520;
  if (IALTH > 1) goto L700;
  if (L == LMAX) goto L700;
  for (int i = 1; i <= N; ++i) YHc(i, LMAX) = ACOR[i - 1];
  goto L700;
L500:
  KFLAG = KFLAG - 1;
  ++n_efail;
  TN = TOLD;
  I1 = NQNYH + 1;
  for (int JB = 1; JB <= NQ; ++JB) {
    I1 = I1 - NYH;
    for (int i = I1; i <= NQNYH; ++i) YH1[i] = YH1[i] - YH1[i + NYH];
  }
  RMAX = 2.0;
  if (std::fabs(H) <= HMIN * 1.00001) goto L660;
  if (KFLAG <= -3) goto L640;
  IREDO = 2;
  RHUP = 0.0;
  goto L540;
L520:
  RHUP = 0.0;
  if (L == LMAX) goto L540;
  for (int i = 1; i <= N; ++i) SAVF[i - 1] = ACOR[i - 1] - YHc(i, LMAX);
  DUP = dvnorm(SAVF.data(), EWT.data()) / TESCO[3][NQ];
  EXUP = 1.0 / (L + 1);
  RHUP = 1.0 / (1.4 * std::pow(DUP, EXUP) + 0.0000014);
L540:
  EXSM = 1.0 / L;
  RHSM = 1.0 / (1.2 * std::pow(DSM, EXSM) + 0.0000012);
  RHDN = 0.0;
  if (NQ == 1) goto L560;
  DDN = dvnorm(&YHc(1, L), EWT.data()) / TESCO[1][NQ];
  EXDN = 1.0 / NQ;
  RHDN = 1.0 / (1.3 * std::pow(DDN, EXDN) + 0.0000013);
L560:
  if (RHSM >= RHUP) goto L570;
  if (RHUP > RHDN) goto L590;
  goto L580;
L570:
  if (RHSM < RHDN) goto L580;
  NEWQ = NQ;
  RH = RHSM;
  goto L620;
L580:
  NEWQ = NQ - 1;
  RH = RHDN;
  if (KFLAG < 0 && RH > 1.0) RH = 1.0;
  goto L620;
L590:
  NEWQ = L;
  RH = RHUP;
  if (RH < 1.1) goto L610;
  R = EL[L] / L;
  for (int i = 1; i <= N; ++i) YHc(i, NEWQ + 1) = ACOR[i - 1] * R;
  goto L630;
L610:
  IALTH = 3;
  goto L700;
L620:
  if (KFLAG == 0 && RH < 1.1) goto L610;
  if (KFLAG <= -2) RH = std::min(RH, 0.2);
  if (NEWQ == NQ) goto L170;
L630:
  NQ = NEWQ;
  L = NQ + 1;
  IRET = 2;
  goto L150;
L640:
  if (KFLAG == -10) goto L660;
  RH = 0.1;
  RH = std::max(HMIN / std::fabs(H), RH);
  H = H * RH;
  for (int i = 1; i <= N; ++i) Y[i - 1] = YHc(i, 1);
  f(TN, Y, SAVF.data());
  NFE = NFE + 1;
  for (int i = 1; i <= N; ++i) YHc(i, 2) = H * SAVF[i - 1];
  IPUP = MITER;
  IALTH = 5;
  if (NQ == 1) goto L200;
  NQ = 1;
  L = 2;
  IRET = 3;
  goto L150;
L660:
  KFLAG = -1;
  goto L720;
L670:
  KFLAG = -2;
  goto L720;
L680:
  KFLAG = -3;
  goto L720;
L690:
  RMAX = 10.0;
L700:
  R = 1.0 / TESCO[2][NQU];
  for (int i = 0; i < N; ++i) ACOR[i] = ACOR[i] * R;
L720:
  HOLD = H;
  JSTART = 1;
}

int Lsodes::call(double* Y, double* T, double TOUT, const double* RTOL, const double* ATOL,
                 int ITASK, int ISTATE, int maxord, int mxstep, double hmax, double tcrit) {
  double H0 = 0, TOLSF, HMX, TNEXT, TOL, SUM, W0, TDIST, RH;
  auto fail = [&](int code) { return code; };
  if (ISTATE < 1 || ISTATE > 3) return -3;
  if (ITASK != 1 && ITASK != 4) return -3;
  if (ISTATE == 1) {
    INIT = 0;
    if (TOUT == *T) return ISTATE;
  } else {
    if (INIT == 0) return -3;
    if (ISTATE == 2) goto L200;
  }
  // Block B
  N = n;
  METH = 2; MITER = 1;
  MAXORD = maxord; if (MAXORD == 0) MAXORD = 100; MAXORD = std::min(MAXORD, 5);
  if (MAXORD < 0) return -3;
  MXSTEP = mxstep; if (MXSTEP < 0) return -3; if (MXSTEP == 0) MXSTEP = 500;
  MXHNIL = 1;
  if (ISTATE == 1) H0 = 0.0;   // RWORK(5) = 0 in the reference's setup
  if (hmax < 0.0) return -3;
  HMXI = 0.0; if (hmax > 0.0) HMXI = 1.0 / hmax;
  HMIN = 0.0;
  for (int i = 0; i < N; ++i) if (RTOL[i] < 0.0 || ATOL[i] < 0.0) return -3;
  if (!analysed) {
    lu.analyse(N, ia, ja);
    analysed = true;
    Pval.assign(ja.size(), 0.0); Jval.assign(ja.size(), 0.0);
  }
  if (ISTATE == 1) {
    NYH = N;
    YH.assign((size_t)N * 7, 0.0); EWT.assign(N, 0.0); SAVF.assign(N, 0.0); ACOR.assign(N, 0.0);
    FTEM.assign(N, 0.0);
  }
  // DIPREP/DPREP redo the sparse preprocessing on ISTATE = 1 and 3; the symbolic
  // result is identical every time, but the saved P is zeroed (src/opkda1.f:1493-1494)
  std::fill(Pval.begin(), Pval.end(), 0.0);
  Padd = 0.0;
  if (ISTATE == 3) {
    JSTART = -1;
    goto L200;
  }
  // Block C (ISTATE = 1)
  TN = *T;
  NST = 0;
  H = 1.0;
  for (int i = 0; i < N; ++i) YH[i] = Y[i];
  f(*T, Y, &YH[NYH]);
  NFE = 1;
  dewset(RTOL, ATOL);
  if (!ewt_invert_ok()) return -3;
  if (ITASK == 4) {
    TCRIT = tcrit;
    if ((TCRIT - TOUT) * (TOUT - *T) < 0.0) return -3;
    if (H0 != 0.0 && (*T + H0 - TCRIT) * H0 > 0.0) H0 = TCRIT - *T;
  }
  UROUND = std::numeric_limits<double>::epsilon();
  JSTART = 0;
  MSBJ = 50; NSLJ = 0; CCMXJ = 0.2; PSMALL = 1000.0 * UROUND; RBIG = 0.01 / PSMALL;
  NHNIL = 0; NJE = 0; NLU = 0; NSLAST = 0; HU = 0.0; NQU = 0;
  CCMAX = 0.3; MAXCOR = 3; MSBP = 20; MXNCF = 10;
  iplost = 0; CON0 = 0; CONMIN = 0;
  if (H0 == 0.0) {
    TDIST = std::fabs(TOUT - *T);
    W0 = std::max(std::fabs(*T), std::fabs(TOUT));
    if (TDIST < 2.0 * UROUND * W0) return -3;
    TOL = RTOL[0];
    for (int i = 0; i < N; ++i) TOL = std::max(TOL, RTOL[i]);
    if (TOL <= 0.0) {
      for (int i = 0; i < N; ++i) {
        double AYI = std::fabs(Y[i]);
        if (AYI != 0.0) TOL = std::max(TOL, ATOL[i] / AYI);
      }
    }
    TOL = std::max(TOL, 100.0 * UROUND);
    TOL = std::min(TOL, 0.001);
    SUM = dvnorm(&YH[NYH], EWT.data());
    SUM = 1.0 / (TOL * W0 * W0) + TOL * SUM * SUM;
    H0 = 1.0 / std::sqrt(SUM);
    H0 = std::min(H0, TDIST);
    H0 = std::copysign(H0, TOUT - *T);
  }
  RH = std::fabs(H0) * HMXI;
  if (RH > 1.0) H0 = H0 / RH;
  H = H0;
  for (int i = 0; i < N; ++i) YH[NYH + i] = H0 * YH[NYH + i];
  goto L270;
  // Block D
L200:
  NSLAST = NST;
  if (ITASK == 1) {
    if ((TN - TOUT) * H < 0.0) goto L250;
    dintdy(TOUT, Y);
    *T = TOUT;
    goto L420;
  }
  // ITASK = 4, label 230
  TCRIT = tcrit;
  if ((TN - TCRIT) * H > 0.0) return -3;
  if ((TCRIT - TOUT) * H < 0.0) return -3;
  if ((TN - TOUT) * H < 0.0) goto L245;
  dintdy(TOUT, Y);
  *T = TOUT;
  goto L420;
L245:
  HMX = std::fabs(TN) + std::fabs(H);
  IHIT = std::fabs(TN - TCRIT) <= 100.0 * UROUND * HMX;
  if (IHIT) goto L400;
  TNEXT = TN + H * (1.0 + 4.0 * UROUND);
  if ((TNEXT - TCRIT) * H <= 0.0) goto L250;
  H = (TCRIT - TN) * (1.0 - 4.0 * UROUND);
  if (ISTATE == 2) JSTART = -2;
  // Block E
L250:
  if ((NST - NSLAST) >= MXSTEP) { ISTATE = -1; goto L580; }
  dewset(RTOL, ATOL);
  if (!ewt_invert_ok()) { ISTATE = -6; goto L580; }
L270:
  TOLSF = UROUND * dvnorm(YH.data(), EWT.data());
  if (TOLSF > 1.0) {
    TOLSF = TOLSF * 2.0;
    if (NST == 0) return -3;
    ISTATE = -2;
    goto L580;
  }
  if ((TN + H) == TN) NHNIL = NHNIL + 1;
  dstode(Y);
  switch (1 - KFLAG) {
    case 1: break;
    case 2: ISTATE = -4; goto L560;
    case 3: ISTATE = -5; goto L560;
    default: ISTATE = -7; goto L580;
  }
  // Block F
  INIT = 1;
  if (ITASK == 1) {
    if ((TN - TOUT) * H < 0.0) goto L250;
    dintdy(TOUT, Y);
    *T = TOUT;
    goto L420;
  }
  // ITASK = 4, label 340
  if ((TN - TOUT) * H < 0.0) goto L345;
  dintdy(TOUT, Y);
  *T = TOUT;
  goto L420;
L345:
  HMX = std::fabs(TN) + std::fabs(H);
  IHIT = std::fabs(TN - TCRIT) <= 100.0 * UROUND * HMX;
  if (IHIT) goto L400;
  TNEXT = TN + H * (1.0 + 4.0 * UROUND);
  if ((TNEXT - TCRIT) * H <= 0.0) goto L250;
  H = (TCRIT - TN) * (1.0 - 4.0 * UROUND);
  JSTART = -2;
  goto L250;
  // Block G
L400:
  for (int i = 0; i < N; ++i) Y[i] = YH[i];
  *T = TN;
  if (ITASK == 4 && IHIT) *T = TCRIT;
L420:
  return 2;
  // Block H
L560: {
    double BIG = 0.0;
    IMXER = 1;
    for (int i = 0; i < N; ++i) {
      double SIZE = std::fabs(ACOR[i] * EWT[i]);
      if (BIG >= SIZE) continue;
      BIG = SIZE;
      IMXER = i + 1;
    }
  }
L580:
  for (int i = 0; i < N; ++i) Y[i] = YH[i];
  *T = TN;
  (void)fail;
  return ISTATE;
}

}  // namespace raco

using namespace raco;
extern "C" {

raco_lsodes* raco_lsodes_create(int neq, const int* ia, const int* ja, raco_f_cb f, raco_jac_cb jac,
                                void* ctx) {
  raco_lsodes* h = new raco_lsodes();
  Lsodes& s = h->s;
  s.n = neq;
  s.ia.assign(ia, ia + neq + 1);
  s.ja.assign(ja, ja + (ia[neq] - 1));
  s.f = [=](double t, const double* y, double* ydot) { f(neq, t, y, ydot, ctx); };
  s.jac_col = [=](double t, const double* y, int j, double* pdj) { jac(neq, t, y, j, pdj, ctx); };
  return h;
}
void raco_lsodes_free(raco_lsodes* h) { delete h; }
int raco_lsodes_call(raco_lsodes* h, double* y, double* t, double tout, const double* rtol,
                     const double* atol, int itask, int istate, int maxord, int mxstep, double hmax,
                     double tcrit) {
  return h->s.call(y, t, tout, rtol, atol, itask, istate, maxord, mxstep, hmax, tcrit);
}
void raco_lsodes_stats(const raco_lsodes* h, int* out, double* hu) {
  const Lsodes& s = h->s;
  out[0] = s.NST; out[1] = s.NFE; out[2] = s.NJE; out[3] = s.NQU; out[4] = s.NQ; out[5] = s.IMXER;
  out[6] = s.lu.nnz_a; out[7] = s.NLU; out[8] = s.lu.nzl; out[9] = s.lu.nzu;
  if (hu) *hu = s.HU;
}

}  // extern "C"
