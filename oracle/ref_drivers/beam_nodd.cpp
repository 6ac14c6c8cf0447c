// beam_nodd.cpp -- reference-built driver (oracle/_ref/beam_nodd).
// TEST / ORACLE / CPU-BASELINE INFRASTRUCTURE, not product code.
//
// Builds the BEAM example without domain decomposition exactly the way
// examples/BEAM.h:403-410 (BEAM::SOLVE_NODD) does -- MESH_NODD, TRANSFER,
// STIF_MATR, CONSTRAINT(1) -- with the reference's own headers compiled from
// where they lie (/root/reference), then
//   * dumps the multigrid hierarchy (mgpi.consStif[l], mgpi.realProl[l]), the
//     right-hand side consForc,
//   * runs the reference MGPIS::CG_SOLV / MGPIS::MULT_VCYC on it and dumps
//     the results (golden vectors), with wall-clock timings.
//
// usage: beam_nodd --glob G [--divi a,b,c] [--out file.ddpk] [--nomat]
//                  [--solve 0|1] [--jacobi 0|1] [--extra 0|1] [--reps R]
//                  [--bench-steps K --bench-warmup W]   (reference arm of bench.py: W untimed +
//                   K timed MGPIS::CG_SOLV(1,...) calls, wall clock around the K calls)
#include "examples/BEAM.h"
#include "ddpk_io.h"
#include "ref_capture.h"

int main(int argc, char **argv) {
	long glob = 2, doSolve = 1, doJacobi = 0, reps = 1, noMat = 0, benchSteps = 0, benchWarm = 0, doExtra = 0;
	std::vector<long> divi;
	std::string out;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--glob") glob = std::stol(next());
		else if (a == "--divi") {
			std::stringstream ss(next()); std::string t;
			while (std::getline(ss, t, ',')) divi.push_back(std::stol(t));
		}
		else if (a == "--out") out = next();
		else if (a == "--solve") doSolve = std::stol(next());
		else if (a == "--jacobi") doJacobi = std::stol(next());
		else if (a == "--reps") reps = std::stol(next());
		else if (a == "--nomat") noMat = 1;
		else if (a == "--extra") doExtra = std::stol(next());
		else if (a == "--bench-steps") benchSteps = std::stol(next());
		else if (a == "--bench-warmup") benchWarm = std::stol(next());
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;           // keep the reference's chatter off stdout
	BEAM beam(0);
	beam.globLeve = glob;
	if (divi.size() == 3) beam.diviNumb = divi;
	beam.MESH_NODD(0);
	MULTIGRID &mg = beam.multGrid[0];
	mg.TRANSFER();
	mg.STIF_MATR();
	mg.CONSTRAINT(1);
	double tSetup = now_s() - t0;
	MGPIS &mgpi = mg.mgpi;
	long L = mgpi.maxiLeve;
	long n = mgpi.consStif[L].rows();

	DDPK_WRITER *w = out.empty() ? nullptr : new DDPK_WRITER(out);
	if (w) {
		w->scalar_i64("maxiLeve", L);
		for (long l = 0; l <= L; l++) {
			if (!noMat) w->csr("consStif" + std::to_string(l), mgpi.consStif[l]);
			else { long s[3] = {mgpi.consStif[l].rows(), mgpi.consStif[l].cols(), mgpi.consStif[l].nonZeros()};
				w->i64("consStif" + std::to_string(l) + ".dims", s, 3); }
		}
		for (long l = 0; l < L; l++) if (!noMat) w->csr("realProl" + std::to_string(l), mgpi.realProl[l]);
		w->vec("consForc", mg.consForc);
	}
	std::ostringstream js;
	js << std::setprecision(17);
	js << "{\"example\":\"BEAM_NODD\",\"globLeve\":" << glob << ",\"levels\":[";
	for (long l = 0; l <= L; l++) js << (l ? "," : "") << "[" << mgpi.consStif[l].rows() << "," << mgpi.consStif[l].nonZeros() << "]";
	js << "],\"setup_s\":" << tSetup;

	if (doSolve) {
		// one V-cycle on b (kernel-level golden for MULT_VCYC, MGPIS.h:55-128)
		{
			DIRE_SOLV direSolv;
			direSolv.compute(mgpi.consStif[0]);
			Eigen::VectorXd z = Eigen::VectorXd::Zero(n);
			mgpi.MULT_VCYC(L, mg.consForc, z, direSolv);
			if (w) w->vec("vcyc_of_consForc", z);
			// coarse solve golden
			Eigen::VectorXd b0 = Eigen::VectorXd::LinSpaced(mgpi.consStif[0].rows(), -1.0, 1.0);
			Eigen::VectorXd x0 = direSolv.solve(b0);
			if (w) { w->vec("coarse_rhs", b0); w->vec("coarse_sol", x0); }
		}
		Eigen::VectorXd x;
		double best = 1e300; long iters = 0;
		for (long r = 0; r < reps; r++) {
			cap.buf.str("");
			double t1 = now_s();
			mgpi.CG_SOLV(1, mg.consForc, x);
			double dt = now_s() - t1;
			best = std::min(best, dt);
			iters = cap.last_iteration_plus1();
		}
		if (w) { w->vec("cg_mg_x", x); w->scalar_i64("cg_mg_iters", iters); }
		double resid = (mg.consForc - mgpi.consStif[L] * x).norm();
		js << ",\"cg_mg_iters\":" << iters << ",\"cg_mg_s\":" << best
		   << ",\"dof_iter_per_s\":" << (double)n * iters / best
		   << ",\"x_norm\":" << x.norm() << ",\"x_maxabs\":" << x.cwiseAbs().maxCoeff()
		   << ",\"true_resid\":" << resid << ",\"b_norm\":" << mg.consForc.norm();
		// expand to full mesh (OUTP_SUB1, MULTIGRID.h:1263-1281)
		Eigen::VectorXd full;
		mg.OUTP_SUB1(x, full);
		if (w) w->vec("outp_sub1_of_x", full);
	}
	if (doJacobi) {
		Eigen::VectorXd x;
		cap.buf.str("");
		double t1 = now_s();
		mgpi.CG_SOLV(0, mg.consForc, x);
		double dt = now_s() - t1;
		long iters = cap.last_iteration_plus1();
		if (w) { w->vec("cg_jacobi_x", x); w->scalar_i64("cg_jacobi_iters", iters); }
		js << ",\"cg_jacobi_iters\":" << iters << ",\"cg_jacobi_s\":" << dt;
	}
	if (doExtra) {
		// the other MGPIS drivers: BiCGSTAB_SOLV (MGPIS.h:350-432) and MULT_SOLV (MGPIS.h:130-160)
		Eigen::VectorXd x;
		cap.buf.str("");
		mgpi.BiCGSTAB_SOLV(1, mg.consForc, x);
		long itb = cap.last_iteration_plus1();
		if (w) { w->vec("bicgstab_mg_x", x); w->scalar_i64("bicgstab_mg_iters", itb); }
		cap.buf.str("");
		mgpi.MULT_SOLV(mg.consForc, x);
		long itm = cap.last_iteration_plus1() - 1;   // MULT_SOLV prints iterNumb itself (:156)
		if (w) { w->vec("mult_solv_x", x); w->scalar_i64("mult_solv_iters", itm); }
		cap.buf.str("");
		mgpi.GMRES_SOLV(1, mg.consForc, x);   // MGPIS.h:227-348 (prints iterNumb itself, :344)
		long itg = cap.last_iteration_plus1() - 1;
		if (w) { w->vec("gmres_mg_x", x); w->scalar_i64("gmres_mg_iters", itg); }
		js << ",\"bicgstab_mg_iters\":" << itb << ",\"mult_solv_iters\":" << itm << ",\"gmres_mg_iters\":" << itg;
	}
	if (benchSteps > 0) {
		Eigen::VectorXd x;
		for (long r = 0; r < benchWarm; r++) mgpi.CG_SOLV(1, mg.consForc, x);
		long iters = 0;
		double t1 = now_s();
		for (long r = 0; r < benchSteps; r++) {
			cap.buf.str("");
			mgpi.CG_SOLV(1, mg.consForc, x);
			iters += cap.last_iteration_plus1();
		}
		double dt = now_s() - t1;
		js << ",\"bench_steps\":" << benchSteps << ",\"bench_warmup\":" << benchWarm
		   << ",\"bench_s\":" << dt << ",\"bench_iters\":" << iters
		   << ",\"bench_dof_iter_per_s\":" << (double)n * iters / dt;
	}
	js << ",\"threads\":1}";
	delete w;
	cap.release();
	std::cout << js.str() << std::endl;
	return 0;
}
