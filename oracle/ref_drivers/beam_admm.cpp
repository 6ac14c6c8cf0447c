// beam_admm.cpp -- reference-built driver (oracle/_ref/beam_admm).  TEST / ORACLE / CPU-BASELINE
// INFRASTRUCTURE.  The reference's BEAM example WITH domain decomposition (examples/BEAM.h:390-401,
// :424-609: ESTA_SURF, MESH_DD, SOLVE_DD -> contact search on the tied subdomain interfaces,
// MCONTACT::ESTABLISH) run unchanged; ADMM_HOOK (admm_hook.h) takes over MCONTACT::CONTACT_ANALYSIS.
// BASELINE.json config "BEAM ... 8 subdomains" = --doma 8,1,1.
//
// usage: beam_admm --glob G --doma a,b,c [--divi a,b,c] [--musc 0|1] [--out f.ddpk] [--ref-iters K|-1] [--nomat]
#include "MCONTACT.h"
#include "admm_hook.h"
#include "examples/BEAM.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/BEAM.cpp:37-38
	omp_set_dynamic(1);
	long glob = 2, musc = 1;
	std::vector<long> doma = {8, 1, 1}, divi;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		auto list = [&](std::vector<long> &v) { v.clear(); std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) v.push_back(std::stol(t)); };
		if (a == "--glob") glob = std::stol(next());
		else if (a == "--musc") musc = std::stol(next());
		else if (a == "--doma") list(doma);
		else if (a == "--divi") list(divi);
		else if (a == "--out") g_admmOpts.out = next();
		else if (a == "--ref-iters") g_admmOpts.refIters = std::stol(next());
		else if (a == "--nomat") g_admmOpts.noMat = true;
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	BEAM beam(1);                         // coloSett = 1: domain decomposition (examples/BEAM.cpp:52-83)
	beam.muscSett = musc;                 // SELE_COSP(): 1 = macroscopic problem, 0 = none
	beam.domaNumb = doma;                 // every diviNumb must be divisible by the matching domaNumb
	beam.doleMcsc.assign(doma[0] * doma[1] * doma[2], 1);
	if (divi.size() == 3) beam.diviNumb = divi;
	beam.globLeve = glob;
	{
		std::ostringstream tl;
		tl << ",\"example\":\"BEAM_DD\",\"globLeve\":" << glob << ",\"domaNumb\":[" << doma[0] << "," << doma[1] << "," << doma[2] << "]";
		g_admmOpts.jsonTail = tl.str();
	}
	beam.SOLVE();
	cap.release();
	std::string js = g_admmOpts.json;
	js.pop_back();
	std::cout << js << g_admmOpts.jsonTail << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	return 0;
}
