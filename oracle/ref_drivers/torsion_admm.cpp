// torsion_admm.cpp -- reference-built driver (oracle/_ref/torsion_admm).  TEST / ORACLE / CPU-BASELINE
// INFRASTRUCTURE.  The reference's TORSION example with domain decomposition (examples/TORSION.h:408-634:
// ESTA_SURF, MESH_DD, SOLVE_DD -> contact search on the tied subdomain interfaces of the hollow shaft,
// MCONTACT::ESTABLISH) run unchanged; ADMM_HOOK (admm_hook.h) takes over MCONTACT::CONTACT_ANALYSIS.
// The example's menus (examples/TORSION.cpp:52-83) are domaNumb 1x8x4 / 1x16x4 / 1x16x8 at
// globInho=1, globHomo=4; --homo lowers the synthetic refinement.
//
// usage: torsion_admm [--doma a,b,c] [--divi a,b,c] [--inho I] [--homo H] [--musc 0..3] [--dole D]
//                     [--out f.ddpk] [--ref-iters K|-1] [--nomat]
#include "MCONTACT.h"
#include "admm_hook.h"
#include "examples/TORSION.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/TORSION.cpp:37-38
	omp_set_dynamic(1);
	long inho = 1, homo = 2, musc = 2, dole = 2;
	std::vector<long> doma = {1, 8, 4}, divi;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		auto list = [&](std::vector<long> &v) { v.clear(); std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) v.push_back(std::stol(t)); };
		if (a == "--inho") inho = std::stol(next());
		else if (a == "--homo") homo = std::stol(next());
		else if (a == "--musc") musc = std::stol(next());
		else if (a == "--dole") dole = std::stol(next());
		else if (a == "--doma") list(doma);
		else if (a == "--divi") list(divi);
		else if (a == "--out") g_admmOpts.out = next();
		else if (a == "--ref-iters") g_admmOpts.refIters = std::stol(next());
		else if (a == "--nomat") g_admmOpts.noMat = true;
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	TORSION tors(1);                      // coloSett = 1: domain decomposition (examples/TORSION.cpp:53-58)
	tors.muscSett = musc;                 // SELE_COSP(): bit 0 macroscopic problem, bit 1 interface-eliminated problem
	tors.domaNumb = doma;                 // every diviNumb must be divisible by the matching domaNumb
	tors.doleMcsc.assign(doma[0] * doma[1] * doma[2], dole);
	if (divi.size() == 3) tors.diviNumb = divi;
	tors.globInho = inho;
	tors.globHomo = homo;
	{
		std::ostringstream tl;
		tl << ",\"example\":\"TORSION_DD\",\"globInho\":" << inho << ",\"globHomo\":" << homo << ",\"domaNumb\":[" << doma[0] << "," << doma[1] << "," << doma[2] << "]"
		   << ",\"analytic_u\":1.159111630361142e-06";   // examples/TORSION.h:49
		g_admmOpts.jsonTail = tl.str();
	}
	tors.SOLVE();
	cap.release();
	std::string js = g_admmOpts.json;
	js.pop_back();
	std::cout << js << g_admmOpts.jsonTail << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	return 0;
}
