// cylinder_admm.cpp -- reference-built driver (oracle/_ref/cylinder_admm).  TEST / ORACLE / CPU-BASELINE
// INFRASTRUCTURE.  The reference's CYLINDER example (Hertzian contact of stacked cylinders,
// examples/CYLINDER.h:171-853: MESH with local refinement towards the contact bands -> hanging nodes,
// contact search, MCONTACT::ESTABLISH) run unchanged; ADMM_HOOK takes over MCONTACT::CONTACT_ANALYSIS.
// The example's menus (examples/CYLINDER.cpp:53-81) are copyNumb 4 / 8 / 16 (32 / 64 / 128 subdomains) at
// globInho=3, globHomo=0, locaLeve=7; --copy/--loca give the reduced variants (copyNumb must divide 16).
//
// usage: cylinder_admm [--copy C] [--inho I] [--homo H] [--loca L] [--musc 0|1] [--out f.ddpk] [--ref-iters K|-1] [--nomat]
#include "MCONTACT.h"
#include "admm_hook.h"
#include "examples/CYLINDER.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/CYLINDER.cpp:39-40
	omp_set_dynamic(1);
	long copy = 1, inho = 3, homo = 0, loca = 2, musc = 1;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--copy") copy = std::stol(next());
		else if (a == "--inho") inho = std::stol(next());
		else if (a == "--homo") homo = std::stol(next());
		else if (a == "--loca") loca = std::stol(next());
		else if (a == "--musc") musc = std::stol(next());
		else if (a == "--out") g_admmOpts.out = next();
		else if (a == "--ref-iters") g_admmOpts.refIters = std::stol(next());
		else if (a == "--nomat") g_admmOpts.noMat = true;
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	CYLINDER cyli;
	cyli.copyNumb = copy;                 // examples/CYLINDER.cpp:55
	cyli.muscSett = musc;                 // SELE_COSP_1(): 1 = macroscopic problem, 0 = none
	cyli.globInho = inho;
	cyli.globHomo = homo;
	cyli.locaLeve = loca;
	{
		std::ostringstream tl;
		tl << ",\"example\":\"CYLINDER\",\"copyNumb\":" << copy << ",\"globInho\":" << inho << ",\"globHomo\":" << homo << ",\"locaLeve\":" << loca;
		g_admmOpts.jsonTail = tl.str();
	}
	cyli.SOLVE();
	cap.release();
	std::string js = g_admmOpts.json;
	js.pop_back();
	std::cout << js << g_admmOpts.jsonTail << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	return 0;
}
