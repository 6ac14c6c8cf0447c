// admm_hook.h -- TEST / ORACLE INFRASTRUCTURE.  Include AFTER the reference's MCONTACT.h
// and BEFORE an examples/*.h header: every call `CONTACT_ANALYSIS()` written inside the
// example's SOLVE() (e.g. examples/BLOCK.h:707) then lands in ADMM_HOOK(*this) instead,
// right after MCONTACT::ESTABLISH() has built all operators -- the upload point named in
// SURVEY.md §3.4.  The hook
//   1. dumps every operator the ADMM loop consumes (SURVEY.md §8a rows a6-a13 and the
//      macroscopic-problem operators of MCONTACT.h:2540-2573) into a DDPK file,
//   2. optionally runs the UNTOUCHED reference loop MCONTACT::CONTACT_ANALYSIS
//      (MCONTACT.h:2493-2723), either to convergence or for exactly the first K iterations
//      (the loop announces every iteration on std::cout, MCONTACT.h:2505; a stream buffer
//      installed by the hook counts the announcements and, on the one that opens iteration K,
//      i.e. after K complete passes of the loop body incl. MONITOR, dumps the state and leaves
//      the process from the main thread -- a deterministic stop, no second thread),
//      and dumps its results (resuMoni.txt rows, resuDisp, inteAuxi, inteLagr,
//      resuCont_*.txt) as golden vectors.
#ifndef ADMM_HOOK_H
#define ADMM_HOOK_H
#include <unistd.h>
#include <functional>
#include <mutex>
#include <omp.h>
#include "ddpk_io.h"
#include "ref_capture.h"

struct ADMM_HOOK_OPTS {
	std::string out;        // DDPK output ("" = none)
	long refIters = 0;      // 0: run the reference loop to convergence; K>0: first K iterations only; <0: skip
	bool noMat = false;     // do not dump multigrid operators (sizes only)
	std::string json;       // filled by the hook: one JSON object describing the run
	std::string jsonTail;   // set by the driver: extra ,"key":value pairs
} g_admmOpts;

static std::vector<std::vector<double>> READ_TABLE(const std::string &path) {
	std::vector<std::vector<double>> rows;
	std::ifstream f(path);
	std::string line;
	while (std::getline(f, line)) {
		std::stringstream ss(line);
		std::vector<double> r; double v;
		while (ss >> v) r.push_back(v);
		if (!r.empty()) rows.push_back(r);
	}
	return rows;
}

static void DUMP_TABLE(DDPK_WRITER &w, const std::string &name, const std::vector<std::vector<double>> &rows) {
	std::vector<double> flat;
	long ncol = rows.empty() ? 0 : (long)rows[0].size();
	long nrow = 0;
	for (auto &r : rows) if ((long)r.size() == ncol) { flat.insert(flat.end(), r.begin(), r.end()); nrow++; }
	long shape[2] = {nrow, ncol};
	w.i64(name + ".shape", shape, 2);
	w.f64(name, flat.data(), flat.size());
}

// A factorised DIRE_SOLV (Eigen::SimplicialLDLT, PREP.h:107) as plain arrays: fill-reducing
// permutation p (solve does y[p[i]] = b[i], SimplicialCholesky.h:148-171), strictly-lower unit
// factor L in RowMajor CSR and the diagonal D.  x = P^T L^-T D^-1 L^-1 P b.
static void DUMP_LDLT(DDPK_WRITER &w, const std::string &name, const DIRE_SOLV &sol) {
	if (sol.info() != Eigen::Success || sol.rows() == 0) return;
	Eigen::SparseMatrix<double, Eigen::RowMajor> L = sol.matrixL().nestedExpression();
	w.csr(name + ".L", L);
	Eigen::VectorXd D = sol.vectorD();
	w.vec(name + ".D", D);
	std::vector<int> p(sol.rows());
	for (long i = 0; i < sol.rows(); i++) p[i] = sol.permutationP().indices()(i);
	w.i32(name + ".perm", p.data(), p.size());
}

// Watches the reference's std::cout traffic during the loop (every write takes a mutex: the reference prints from
// its OpenMP threads):
//   * "The <tc>-th iteration" (MCONTACT.h:2505, main thread) -> time stamp; announcement number `stopAt` (the one
//     that opens iteration stopAt, i.e. after stopAt complete passes incl. MONITOR) triggers `onStop`;
//   * "#Iteration: <k>" (MGPIS.h:221, one per MGPIS::CG_SOLV call, printed by the thread that ran the solve):
//     iterNumb = k + 1 is added to the CG iteration total of the current ADMM iteration.  The two tokens arrive
//     as consecutive writes of ONE thread, so a thread-local flag pairs them even when threads interleave.
class LOOP_WATCH_BUF : public std::streambuf {
public:
	LOOP_WATCH_BUF(std::streambuf *next, long stopAt, std::function<void()> onStop) : next_(next), stopAt_(stopAt), onStop_(onStop) {}
	std::vector<double> iterStart;      // time stamp of every announcement
	std::vector<long> cgIters, cgCalls; // per ADMM iteration: sum of iterNumb over the CG_SOLV calls, number of calls
protected:
	std::streamsize xsputn(const char *s, std::streamsize n) override {
		static thread_local bool expectCount = false;
		std::string chunk(s, (size_t)n);
		std::lock_guard<std::mutex> g(mtx_);
		if (expectCount) {
			expectCount = false;
			char *end = nullptr;
			long k = std::strtol(chunk.c_str(), &end, 10);
			if (end != chunk.c_str() && !cgIters.empty()) { cgIters.back() += k + 1; cgCalls.back() += 1; }
		}
		if (chunk.find("#Iteration: ") != std::string::npos) expectCount = true;
		if (chunk.find("-th iteration") != std::string::npos) {
			if (stopAt_ >= 0 && (long)iterStart.size() == stopAt_) { iterStart.push_back(now_s()); onStop_(); }
			iterStart.push_back(now_s());
			cgIters.push_back(0);
			cgCalls.push_back(0);
		}
		return next_ ? next_->sputn(s, n) : n;
	}
	int overflow(int c) override { return (next_ && c != EOF) ? next_->sputc((char)c) : c; }
	int sync() override { return next_ ? next_->pubsync() : 0; }
private:
	std::mutex mtx_;
	std::streambuf *next_;
	long stopAt_;
	std::function<void()> onStop_;
};

inline long ADMM_HOOK(MCONTACT &mc) {
	typedef Eigen::SparseMatrix<double, Eigen::RowMajor> SPM;
	const long nb = mc.multGrid.size(), ni = mc.searCont.size();
	DDPK_WRITER *w = g_admmOpts.out.empty() ? nullptr : new DDPK_WRITER(g_admmOpts.out);
	std::ostringstream js;
	js << std::setprecision(17) << "{\"bodies\":" << nb << ",\"interfaces\":" << ni << ",\"muscSett\":" << mc.muscSett;
	js << ",\"body_dof\":[";
	for (long v = 0; v < nb; v++) js << (v ? "," : "") << mc.multGrid[v].mgpi.consStif[mc.multGrid[v].mgpi.maxiLeve].rows();
	js << "]";
	if (w) {
		w->scalar_i64("nbody", nb); w->scalar_i64("niface", ni); w->scalar_i64("muscSett", mc.muscSett);
		for (long v = 0; v < nb; v++) {
			MULTIGRID &mg = mc.multGrid[v];
			const long L = mg.mgpi.maxiLeve;
			std::string p = "body" + std::to_string(v) + ".";
			w->scalar_i64(p + "maxiLeve", L);
			w->scalar_i64(p + "nfull", 3 * (long)mg.nodeCoor.size());
			w->scalar_i64(p + "doleMcsc", mc.doleMcsc.size() > (size_t)v ? mc.doleMcsc[v] : 0);
			if (!g_admmOpts.noMat) {
				for (long l = 0; l <= L; l++) w->csr(p + "consStif" + std::to_string(l), mg.mgpi.consStif[l]);
				for (long l = 0; l < L; l++) w->csr(p + "realProl" + std::to_string(l), mg.mgpi.realProl[l]);
			}
			w->vec(p + "consForc", mg.consForc);
			{   // node coordinates by node id (analytic checks, e.g. examples/TORSION.h:49)
				std::vector<double> xyz(3 * mg.nodeCoor.size(), 0.0);
				for (const auto &it : mg.nodeCoor)
					if (it.first >= 0 && 3 * (size_t)it.first + 2 < xyz.size()) for (int k = 0; k < 3; k++) xyz[3 * it.first + k] = it.second[k];
				w->f64(p + "nodeCoor", xyz.data(), xyz.size());
			}
			// ADDITIONAL_FORCE (MULTIGRID.h:1257-1261) as ONE operator: n_L x 3 n_nodes
			SPM forcOper = mg.consOper[L] * SPM(mg.prolOper[L].transpose()) * SPM(mg.earlTran.transpose());
			w->csr(p + "forcOper", forcOper);
			// OUTP_SUB1 (MULTIGRID.h:1263-1281) = forcOper^T * u + dispCons
			Eigen::VectorXd zero = Eigen::VectorXd::Zero(mg.mgpi.consStif[L].rows()), dispCons;
			mg.OUTP_SUB1(zero, dispCons);
			w->vec(p + "dispCons", dispCons);
			{   // self-check of the two restatements against the reference functions
				Eigen::VectorXd u = Eigen::VectorXd::Random(mg.mgpi.consStif[L].rows()), full;
				mg.OUTP_SUB1(u, full);
				double e1 = (full - (SPM(forcOper.transpose()) * u + dispCons)).norm() / full.norm();
				Eigen::VectorXd f = Eigen::VectorXd::Random(3 * mg.nodeCoor.size()), g = f;
				mg.ADDITIONAL_FORCE(g);
				double e2 = (g - forcOper * f).norm() / g.norm();
				if (e1 > 1e-13 || e2 > 1e-13) { std::cerr << "forcOper self-check failed " << e1 << " " << e2 << std::endl; std::exit(3); }
			}
			if ((mc.muscSett & 3) && (long)mc.accuProl.size() > v) w->csr(p + "accuProl", mc.accuProl[v]);
			if ((mc.muscSett & 2) && (long)mc.globTran_D_1.size() > v) w->csr(p + "globTran_D_1", mc.globTran_D_1[v]);   // MCONTACT.h:1868-2055
		}
		for (long ts = 0; ts < ni; ts++) {
			std::string p = "if" + std::to_string(ts) + ".";
			long cb[2] = {mc.contBody[ts][0], mc.contBody[ts][1]};
			w->i64(p + "contBody", cb, 2);
			w->scalar_f64(p + "fricCoef", mc.fricCoef[ts]);
			w->scalar_i64(p + "nip", (long)mc.searCont[ts].intePoin.size());
			Eigen::VectorXd gapTerm = mc.pemaInpo[ts] * mc.inpoNgap[ts];   // MCONTACT.h:2636
			w->vec(p + "gapTerm", gapTerm);
			for (long tv = 0; tv < 2; tv++) {
				std::string q = p + "s" + std::to_string(tv) + ".";
				w->csr(q + "systTran", mc.systTran[ts][tv]);
				w->csr(q + "systTran_pena", mc.systTran_pena[ts][tv]);
				w->csr(q + "inteMass", mc.inteMass[ts][tv]);
				w->csr(q + "inteMass_pena", mc.inteMass_pena[ts][tv]);
				w->csr(q + "inpoLagr", mc.inpoLagr[ts][tv]);
				w->csr(q + "inteInpo", mc.inteInpo[ts][tv]);
				w->csr(q + "pemaInpo_r", mc.pemaInpo_r[ts][tv]);
				if (mc.inteMass[ts][tv].rows() < DIRE_MAXI) {   // MCONTACT.h:837-847
					DUMP_LDLT(*w, q + "inteDiso", mc.inteDiso[ts][tv]);
					DUMP_LDLT(*w, q + "inteDiso_pena", mc.inteDiso_pena[ts][tv]);
				}
				if (mc.muscSett & 1) {
					w->csr(q + "globTran", mc.globTran[ts][tv]);
					w->csr(q + "globTran_pena", mc.globTran_pena[ts][tv]);
					w->csr(q + "globTran_D", mc.globTran_D[ts][tv]);
				}
				if (mc.muscSett & 2) w->csr(q + "globTran_1", mc.globTran_1[ts][tv]);   // MCONTACT.h:2124-2298
			}
		}
		if (mc.muscSett & 3) w->i64("baseReco", mc.baseReco.data(), mc.baseReco.size());
		if (mc.muscSett & 1) {
			w->csr("globCoup", mc.globCoup);
			if (mc.globCoup.rows() < DIRE_MAXI) DUMP_LDLT(*w, "coarSolv_D", mc.coarSolv_D);   // MCONTACT.h:1229-1230
		}
		if (mc.muscSett & 2) {   // interface-eliminated coarse problem, MCONTACT::MULTISCALE_1 (MCONTACT.h:1672-2343)
			w->csr("globCoup_1", mc.globCoup_1);
			if (mc.globCoup_1.rows() < DIRE_MAXI) DUMP_LDLT(*w, "coarSolv_D_1", mc.coarSolv_D_1);   // :1857-1858
			w->vec("globForc_1", mc.globForc_1);
		}
	}
	if (mc.muscSett & 1) js << ",\"globCoup_rows\":" << mc.globCoup.rows();
	// ---- the untouched reference loop ------------------------------------------------------
	auto dump_state = [&]() {
		if (!w) return;
		for (long v = 0; v < nb; v++) w->vec("ref.resuDisp" + std::to_string(v), mc.resuDisp[v]);
		for (long ts = 0; ts < ni; ts++) for (long tv = 0; tv < 2; tv++) {
			std::string q = "ref.if" + std::to_string(ts) + ".s" + std::to_string(tv) + ".";
			w->vec(q + "inteAuxi", mc.inteAuxi[ts][tv]);
			w->vec(q + "inteLagr", mc.inteLagr[ts][tv]);
		}
		DUMP_TABLE(*w, "ref.resuMoni", READ_TABLE(DIRECTORY("resuMoni.txt")));
		for (long ts = 0; ts < ni; ts++)   // written by OUTPUT_PRTR in every iteration (MCONTACT.h:2669): the LAST iteration's content
			if (mc.fricCoef[ts] >= 0.0) DUMP_TABLE(*w, "ref.resuCont" + std::to_string(ts), READ_TABLE(DIRECTORY("resuCont_" + std::to_string(ts) + ".txt")));
	};
	auto timing_json = [&](LOOP_WATCH_BUF &wb, double tEnd) {
		// per ADMM iteration: wall time, CG_SOLV calls and the sum of their iteration counts (iterNumb)
		std::ostringstream o;
		o << std::setprecision(9) << ",\"ref_iter_s\":[";
		for (size_t k = 0; k < wb.cgIters.size(); k++) o << (k ? "," : "") << ((k + 1 < wb.iterStart.size() ? wb.iterStart[k + 1] : tEnd) - wb.iterStart[k]);
		o << "],\"ref_cg_iters\":[";
		for (size_t k = 0; k < wb.cgIters.size(); k++) o << (k ? "," : "") << wb.cgIters[k];
		o << "],\"ref_cg_calls\":[";
		for (size_t k = 0; k < wb.cgCalls.size(); k++) o << (k ? "," : "") << wb.cgCalls[k];
		o << "],\"omp_max_threads\":" << omp_get_max_threads();
		return o.str();
	};
	if (g_admmOpts.refIters == 0) {
		double t0 = now_s();
		std::streambuf *prev0 = std::cout.rdbuf();
		LOOP_WATCH_BUF watch0(prev0, -1, []() {});
		std::cout.rdbuf(&watch0);
		(mc.CONTACT_ANALYSIS)();     // parenthesised: not the function-like macro below
		std::cout.rdbuf(prev0);
		double dt = now_s() - t0;
		js << timing_json(watch0, now_s());
		js << ",\"ref_iterNumbReco\":" << mc.iterNumbReco << ",\"ref_admm_s\":" << dt << ",\"ref_MULT_MAXI\":" << MULT_MAXI;
		js << ",\"ref_disp_norm\":[";
		for (long v = 0; v < nb; v++) js << (v ? "," : "") << mc.resuDisp[v].norm();
		js << "]";
		if (w) w->scalar_i64("ref.iterNumbReco", mc.iterNumbReco);
		dump_state();
	} else if (g_admmOpts.refIters > 0) {
		// exactly the first K iterations: state after K passes of the loop body, K resuMoni rows, and the
		// resuCont_*.txt of iteration K-1.  If the loop converges earlier it simply returns.
		std::remove(DIRECTORY("resuMoni.txt").c_str());
		double t0 = now_s();
		std::string head = js.str();
		std::streambuf *prev = std::cout.rdbuf();
		LOOP_WATCH_BUF *self = nullptr;
		LOOP_WATCH_BUF stopper(prev, g_admmOpts.refIters, [&]() {
			if (w) { w->scalar_i64("ref.first_iters", g_admmOpts.refIters); dump_state(); delete w; }
			std::ostringstream o;
			o << std::setprecision(17) << head << ",\"ref_first_iters\":" << g_admmOpts.refIters << ",\"ref_first_iters_s\":" << now_s() - t0
			  << ",\"ref_MULT_MAXI\":" << MULT_MAXI << timing_json(*self, now_s()) << g_admmOpts.jsonTail << "}";
			std::printf("%s\n", o.str().c_str());
			std::fflush(stdout);
			_exit(0);
		});
		self = &stopper;
		std::cout.rdbuf(&stopper);
		(mc.CONTACT_ANALYSIS)();
		std::cout.rdbuf(prev);
		js << ",\"ref_iterNumbReco\":" << mc.iterNumbReco << ",\"ref_admm_s\":" << now_s() - t0 << timing_json(stopper, now_s());   // converged before K iterations
		if (w) w->scalar_i64("ref.iterNumbReco", mc.iterNumbReco);
		dump_state();
	}
	js << "}";
	g_admmOpts.json = js.str();
	delete w;
	return 1;
}

#define CONTACT_ANALYSIS() ADMM_HOOK(*this)
#endif
