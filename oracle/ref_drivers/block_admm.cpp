// block_admm.cpp -- reference-built driver (oracle/_ref/block_admm).  TEST / ORACLE /
// CPU-BASELINE INFRASTRUCTURE.  Runs the reference's BLOCK example (examples/BLOCK.h:510-722:
// MESH, contact search, MCONTACT::ESTABLISH) unchanged, then hands over to ADMM_HOOK
// (admm_hook.h) in place of MCONTACT::CONTACT_ANALYSIS.
//
// usage: block_admm --glob G [--doma a,b,c] [--divi a,b,c] [--musc 0|1] [--out f.ddpk] [--ref-iters K|-1] [--nomat]
#include "MCONTACT.h"
#include "admm_hook.h"
#include "examples/BLOCK.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/BLOCK.cpp:39-40
	omp_set_dynamic(1);
	long glob = 2, musc = 1;
	std::vector<long> doma = {1, 1, 1}, divi;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--glob") glob = std::stol(next());
		else if (a == "--musc") musc = std::stol(next());
		else if (a == "--doma") { doma.clear(); std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) doma.push_back(std::stol(t)); }
		else if (a == "--divi") { std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) divi.push_back(std::stol(t)); }
		else if (a == "--out") g_admmOpts.out = next();
		else if (a == "--ref-iters") g_admmOpts.refIters = std::stol(next());
		else if (a == "--nomat") g_admmOpts.noMat = true;
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	BLOCK bloc;
	bloc.muscSett = musc;                 // examples/BLOCK.cpp:55-64 (menu 0/1/2 + SELE_COSP_1)
	bloc.domaNumb = doma;
	bloc.doleMcsc.assign(3 * doma[0] * doma[1] * doma[2] + 6, 1);
	if (divi.size() == 3) bloc.diviNumb = divi;   // coarsest mesh divisions (default 6,6,6)
	bloc.globLeve = glob;                 // synthetic refinement knob; constructor derives these
	for (long tb = 0; tb < 3; tb++) {     // from globLeve (examples/BLOCK.h:48-53)
		bloc.deltZlen[tb] = bloc.leng[tb] / (bloc.diviNumb[tb] * (1 << bloc.globLeve));
		bloc.uppeZlen[tb] = bloc.leng[tb] - bloc.deltZlen[tb];
	}
	{
		std::ostringstream tl;
		tl << ",\"example\":\"BLOCK\",\"globLeve\":" << glob << ",\"domaNumb\":[" << doma[0] << "," << doma[1] << "," << doma[2] << "]";
		g_admmOpts.jsonTail = tl.str();
	}
	bloc.SOLVE();
	cap.release();
	std::string js = g_admmOpts.json;
	js.pop_back();
	std::cout << js << g_admmOpts.jsonTail << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	return 0;
}
