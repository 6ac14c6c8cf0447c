// lagrange_tap.h -- TEST / ORACLE INFRASTRUCTURE shared by the *_lagrange.cpp drivers.  Include FIRST (before
// MCONTACT.h).  One driver source gives two builds:
//   oracle/_ref/<x>_lagrange                        the UNTOUCHED reference (oracle/Makefile)
//   ddpca-admm_b200/host/_bin/<x>_lagrange_b200     the same with class MGPIS swapped for the B200 overlay
//                                                   (host/Makefile: -DMGPIS=MGPIS_BASE -include host/MGPIS.h)
// and LAGR_REPORT() turns the captured log of MCONTACT::LAGRANGE(1) (MCONTACT.h:2847-3701) into JSON fields.
#ifndef LAGRANGE_TAP_H
#define LAGRANGE_TAP_H
#ifndef _MGPIS_H                       // reference build: MGPIS.h not force-included
#include "PREP.h"
#define MGPIS MGPIS_BASE
#include "MGPIS.h"
#undef MGPIS
#define LAGR_IMPL "reference"
#define LAGR_REFERENCE_BUILD 1
#else                                  // overlay build: host/Makefile passes -DMGPIS=MGPIS_BASE ahead of the force-include
#undef MGPIS
#define LAGR_IMPL "b200"
#endif
#include <chrono>
#include "ddpk_io.h"
#include <mutex>
#include <sstream>
static std::string g_lagrOut;
static double g_lagrSkew = 0.0;   // --skew s (reference build, with --out): also dump a non-symmetric variant of the system
static long g_lagrCalls = 0, g_lagrRows = 0;
static std::vector<double> g_lagrEstaSecs, g_lagrSolvSecs;   // per active-set step: ESTABLISH, BiCGSTAB_SOLV wall time
static inline double lagr_now() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// The tap: class MGPIS of this translation unit derives from the class under test (the reference's, or the B200
// overlay's -- renamed MGPIS_BASE while its header is read, member functions compiled as they are) and wraps the two
// calls MCONTACT::LAGRANGE makes on its local hierarchy (MCONTACT.h:3561-3562) with a timer; in the reference build
// the first call can also be dumped as a fixture.  The per-body `mgpi.ESTABLISH()` calls of MULTIGRID::CONSTRAINT
// (MULTIGRID.h:1251) come first and are not counted: only ESTABLISH calls directly followed by a BiCGSTAB_SOLV are.
class MGPIS : public MGPIS_BASE {
public:
	long ESTABLISH() {
		double t0 = lagr_now();
		long r = MGPIS_BASE::ESTABLISH();
		lastEsta_ = lagr_now() - t0;
		return r;
	}
	long BiCGSTAB_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu) {
		double t0 = lagr_now();
		long r = MGPIS_BASE::BiCGSTAB_SOLV(precSwit, totaForc, resuSolu);
		g_lagrSolvSecs.push_back(lagr_now() - t0);
		g_lagrEstaSecs.push_back(lastEsta_);
		g_lagrRows = consStif[maxiLeve].rows();
		if (g_lagrCalls++ == 0 && !g_lagrOut.empty()) {
			DDPK_WRITER w(g_lagrOut);
			w.scalar_i64("maxiLeve", maxiLeve);
			for (long tl = 0; tl <= maxiLeve; tl++) w.csr("consStif" + std::to_string(tl), consStif[tl]);
			for (long tl = 0; tl < maxiLeve; tl++) w.csr("realProl" + std::to_string(tl), realProl[tl]);
			w.vec("F", totaForc);
			w.vec("U_1", resuSolu);
#ifdef LAGR_REFERENCE_BUILD
			if (g_lagrSkew != 0.0) SKEW_VARIANT(w, totaForc);
#endif
		}
		return r;
	}
#ifdef LAGR_REFERENCE_BUILD
	// A non-symmetric system of the same kind, solved by the UNTOUCHED reference: with sliding friction the condensed
	// matrix of MCONTACT::LAGRANGE loses its symmetry (MCONTACT.h:3376-3413) -- no reduced example reaches that state
	// (BLOCK and CYLINDER_1 are frictionless, DEHW's flanks only touch after hours), so the fixture carries a synthetic
	// one: every strictly-upper entry of K scaled by (1 + s), every strictly-lower one by (1 - s) (symmetric part
	// unchanged, still positive definite), coarse levels by the same Galerkin products as MCONTACT.h:3557-3559, then
	// MGPIS::ESTABLISH, one MULT_VCYC and BiCGSTAB_SOLV(1, F, .) of the reference class.  What it pins: the sweeps with
	// lower != upper^T, and level 0 -- SimplicialLDLT reads only the lower triangle of consStif[0] (PREP.h:107).
	void SKEW_VARIANT(DDPK_WRITER &w, const Eigen::VectorXd &totaForc) {
		typedef Eigen::SparseMatrix<double,Eigen::RowMajor> SPMA;
		MGPIS_BASE skew;
		skew.maxiLeve = maxiLeve;
		skew.realProl = realProl;
		skew.consStif.resize(maxiLeve + 1);
		SPMA fine = consStif[maxiLeve];
		fine.makeCompressed();
		for (long ti = 0; ti < fine.rows(); ti++)
			for (SPMA::InnerIterator iter(fine, ti); iter; ++iter) {
				if (iter.col() > ti) iter.valueRef() = iter.value() * (1.0 + g_lagrSkew);
				else if (iter.col() < ti) iter.valueRef() = iter.value() * (1.0 - g_lagrSkew);
			}
		skew.consStif[maxiLeve] = fine;
		for (long tl = maxiLeve - 1; tl >= 0; tl--)
			skew.consStif[tl] = skew.realProl[tl].transpose() * skew.consStif[tl + 1] * skew.realProl[tl];
		skew.ESTABLISH();
		std::ostringstream text;                       // main thread only: the solver's own log lines
		std::streambuf *prev = std::cout.rdbuf(text.rdbuf());
		Eigen::VectorXd soluSkew;
		skew.BiCGSTAB_SOLV(1, totaForc, soluSkew);
		DIRE_SOLV direSolv;
		direSolv.compute(skew.consStif[0]);
		Eigen::VectorXd vcycSkew = Eigen::VectorXd::Zero(totaForc.rows());
		skew.MULT_VCYC(maxiLeve, totaForc, vcycSkew, direSolv);
		std::cout.rdbuf(prev);
		const std::string log = text.str();
		const size_t p = log.rfind("#Iteration: ");
		w.scalar_f64("skew.s", g_lagrSkew);
		w.scalar_i64("skew.bicgstab_iters", p == std::string::npos ? -1 : std::stol(log.substr(p + 12)) + 1);
		for (long tl = 0; tl < maxiLeve; tl++) w.csr("skew.consStif" + std::to_string(tl), skew.consStif[tl]);
		w.vec("skew.U", soluSkew);
		w.vec("skew.vcyc_of_F", vcycSkew);
	}
#endif
private:
	double lastEsta_ = 0.0;
};
#include "ref_capture.h"

// every BiCGSTAB_SOLV call announces itself (MGPIS.h:351) and ends with "#Iteration: <iterNumb-1>, residual: r/tol"
// + OUTPUT_TIME(":") (MGPIS.h:427-429); the per-iteration lines (MGPIS.h:422) have the same prefix, so the LAST one
// before the next announcement is the call's final line.  "There are <n> unconverged constraints" (MCONTACT.h:3696)
// is the active-set change count of every step but the last.
static std::string LAGR_REPORT(const std::string &log, bool &erro) {
	std::vector<long> iters, changed;
	std::vector<double> resid;
	erro = log.find("ERROR") != std::string::npos;
	for (size_t p = log.find("MGPIS::BiCGSTAB_SOLV (multigrid"); p != std::string::npos;) {
		size_t q = log.find("MGPIS::BiCGSTAB_SOLV (multigrid", p + 1);
		std::string part = log.substr(p, q == std::string::npos ? std::string::npos : q - p);
		size_t r = part.rfind("#Iteration: ");
		if (r != std::string::npos) {
			iters.push_back(std::stol(part.substr(r + 12)) + 1);
			size_t s = part.find("residual: ", r);
			resid.push_back(s == std::string::npos ? -1.0 : std::stod(part.substr(s + 10)));
		}
		size_t u = part.find("There are ");
		if (u != std::string::npos) changed.push_back(std::stol(part.substr(u + 10)));
		p = q;
	}
	std::ostringstream o;
	o << std::setprecision(17) << "\"path\":\"LAGRANGE(1)\",\"impl\":\"" << LAGR_IMPL << "\",\"error\":" << (erro ? "true" : "false")
		<< ",\"converged\":" << (log.find("Converge after ") != std::string::npos ? "true" : "false")
		<< ",\"active_set_steps\":" << iters.size() << ",\"bicgstab_iters\":[";
	for (size_t i = 0; i < iters.size(); i++) o << (i ? "," : "") << iters[i];
	o << "],\"bicgstab_resid\":[";
	for (size_t i = 0; i < resid.size(); i++) o << (i ? "," : "") << resid[i];
	o << "],\"unconverged_constraints\":[";
	for (size_t i = 0; i < changed.size(); i++) o << (i ? "," : "") << changed[i];
	o << "],\"establish_s\":[";
	for (size_t i = 0; i < g_lagrEstaSecs.size(); i++) o << (i ? "," : "") << g_lagrEstaSecs[i];
	o << "],\"bicgstab_s\":[";
	for (size_t i = 0; i < g_lagrSolvSecs.size(); i++) o << (i ? "," : "") << g_lagrSolvSecs[i];
	o << "],\"rows\":" << g_lagrRows;
	return o.str();
}

template <class EXAM> static std::string LAGR_DISP(const EXAM &exam) {
	std::ostringstream o;
	o << std::setprecision(17) << "\"disp_norm\":[";
	for (size_t tv = 0; tv < exam.resuDisp.size(); tv++) o << (tv ? "," : "") << exam.resuDisp[tv].norm();
	o << "],\"disp_size\":[";
	for (size_t tv = 0; tv < exam.resuDisp.size(); tv++) o << (tv ? "," : "") << exam.resuDisp[tv].size();
	o << "]";
	return o.str();
}
#endif
