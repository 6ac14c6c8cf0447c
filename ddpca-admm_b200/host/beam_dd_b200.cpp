// beam_dd_b200.cpp -- the reference's BEAM example with domain decomposition (examples/BEAM.h,
// unchanged) with the hot path on the B200: same set-up as oracle/ref_drivers/beam_admm.cpp, MGPIS is
// the overlay class, MCONTACT::CONTACT_ANALYSIS is DDPCA_CONTACT_ANALYSIS (MCONTACT_B200.h).
//   beam_dd_b200 --glob G --doma a,b,c [--divi a,b,c] [--musc 0|1]
#include "MCONTACT.h"
#define DDPCA_HOOK_CONTACT_ANALYSIS
#include "MCONTACT_B200.h"
#include "examples/BEAM.h"
#include "ref_capture.h"

int main(int argc, char **argv){
	omp_set_nested(1);
	omp_set_dynamic(1);
	long glob = 2, musc = 1;
	std::vector<long> doma = {8, 1, 1}, divi;
	for(int i = 1; i < argc; i ++){
		std::string a = argv[i];
		auto next = [&](){ return std::string(argv[++ i]); };
		auto list = [&](std::vector<long> &v){ v.clear(); std::stringstream ss(next()); std::string t; while(std::getline(ss, t, ',')) v.push_back(std::stol(t)); };
		if(a == "--glob") glob = std::stol(next());
		else if(a == "--musc") musc = std::stol(next());
		else if(a == "--doma") list(doma);
		else if(a == "--divi") list(divi);
		else{ std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	BEAM beam(1);
	beam.muscSett = musc;
	beam.domaNumb = doma;
	beam.doleMcsc.assign(doma[0] * doma[1] * doma[2], 1);
	if(divi.size() == 3) beam.diviNumb = divi;
	beam.globLeve = glob;
	beam.SOLVE();
	cap.release();
	std::string log = cap.buf.str();
	bool erro = log.find("(B200): ERROR") != std::string::npos;
	std::cout << std::setprecision(17) << "{\"example\":\"BEAM_DD\",\"impl\":\"b200\",\"globLeve\":" << glob << ",\"muscSett\":" << musc
		<< ",\"error\":" << (erro ? "true" : "false") << ",\"iterNumbReco\":" << beam.iterNumbReco << ",\"disp_norm\":[";
	for(long tv = 0; tv < beam.resuDisp.size(); tv ++) std::cout << (tv ? "," : "") << beam.resuDisp[tv].norm();
	std::cout << "],\"total_s\":" << now_s() - t0 << "}" << std::endl;
	if(erro){
		size_t p = log.find("(B200): ERROR");
		std::cerr << log.substr(p > 80 ? p - 80 : 0, 400) << std::endl;
	}
	return erro ? 1 : 0;
}
