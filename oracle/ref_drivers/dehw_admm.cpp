// dehw_admm.cpp -- reference-built driver (oracle/_ref/dehw_admm).  TEST / ORACLE / CPU-BASELINE
// INFRASTRUCTURE.  The reference's DEHW example (double-enveloping hourglass worm drive,
// examples/DEHW.h:2217-2289: tooth surfaces, adaptive refinement towards the contact lines, nodal
// rotations, FRICTIONAL contact interfaces mu = 0.08 / 0.2 (:1619), MCONTACT::ESTABLISH) run unchanged;
// ADMM_HOOK takes over MCONTACT::CONTACT_ANALYSIS.  Menus (examples/DEHW.cpp:57-96): --dd 0 = one worm +
// one wheel body, --dd 1 = 34 + 18 subdomains; finest mesh = globInho 1, globHomo 2, locaLeve 3
// (examples/DEHWSURF.h:192-194); --homo/--loca give the reduced variants (globHomo >= 1, locaLeve in 1..3).
//
// usage: dehw_admm [--selo 0|1] [--dd 0|1] [--inho I] [--homo H] [--loca L] [--tape t] [--musc 0|1]
//                  [--out f.ddpk] [--ref-iters K|-1] [--nomat] [--verbose]
#include "MCONTACT.h"
#include "admm_hook.h"
#include "examples/DEHW.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/DEHW.cpp:41-42
	omp_set_dynamic(1);
	long selo = 0, dd = 0, inho = 1, homo = 1, loca = 1, musc = 1;
	double tape = 25.0;
	bool verbose = false;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--selo") selo = std::stol(next());
		else if (a == "--dd") dd = std::stol(next());
		else if (a == "--inho") inho = std::stol(next());
		else if (a == "--homo") homo = std::stol(next());
		else if (a == "--loca") loca = std::stol(next());
		else if (a == "--musc") musc = std::stol(next());
		else if (a == "--tape") tape = std::stod(next());
		else if (a == "--out") g_admmOpts.out = next();
		else if (a == "--ref-iters") g_admmOpts.refIters = std::stol(next());
		else if (a == "--nomat") g_admmOpts.noMat = true;
		else if (a == "--verbose") verbose = true;
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	if (verbose) cap.release();   // show the reference's own progress lines
	tapeCoef = tape;                      // TANG_PEPA(), examples/DEHW.cpp:127-157
	whadCosp = musc;                      // SELE_COSP_1(), examples/DEHW.cpp:185-201
	DEHW solv(1 - selo);                  // ISNO_SELO(): coloSett 1 = driving worm (mu 0.08), 0 = self-locking (mu 0.2)
	solv.dehwSurf.globInho = inho;
	solv.dehwSurf.globHomo = homo;
	solv.dehwSurf.locaLeve = loca;
	{
		std::ostringstream tl;
		tl << ",\"example\":\"DEHW\",\"dd\":" << dd << ",\"selfLocking\":" << selo << ",\"globInho\":" << inho << ",\"globHomo\":" << homo
		   << ",\"locaLeve\":" << loca << ",\"tapeCoef\":" << tape;
		g_admmOpts.jsonTail = tl.str();
	}
	solv.SOLVE(1, dd);
	cap.release();
	std::string js = g_admmOpts.json;
	js.pop_back();
	std::cout << js << g_admmOpts.jsonTail << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	return 0;
}
