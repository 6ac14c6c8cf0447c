// block_b200.cpp -- the reference's BLOCK example (examples/BLOCK.h, unchanged) with the hot path on
// the B200: MGPIS is the overlay class (force-included MGPIS.h), MCONTACT::CONTACT_ANALYSIS is taken
// over by DDPCA_CONTACT_ANALYSIS (MCONTACT_B200.h).  Mesh, contact search and MCONTACT::ESTABLISH run
// as in the reference.  Same command line as oracle/ref_drivers/block_admm.cpp; prints one JSON line
// that can be compared with the pure-reference driver's.
//   block_b200 --glob G [--doma a,b,c] [--divi a,b,c] [--musc 0|1]
#include "MCONTACT.h"
#define DDPCA_HOOK_CONTACT_ANALYSIS
#include "MCONTACT_B200.h"
#include "examples/BLOCK.h"
#include "ref_capture.h"

int main(int argc, char **argv){
	omp_set_nested(1);
	omp_set_dynamic(1);
	long glob = 2, musc = 1;
	std::vector<long> doma = {1, 1, 1}, divi;
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
	BLOCK bloc;
	bloc.muscSett = musc;
	bloc.domaNumb = doma;
	bloc.doleMcsc.assign(3 * doma[0] * doma[1] * doma[2] + 6, 1);
	if(divi.size() == 3) bloc.diviNumb = divi;
	bloc.globLeve = glob;
	for(long tb = 0; tb < 3; tb ++){
		bloc.deltZlen[tb] = bloc.leng[tb] / (bloc.diviNumb[tb] * (1 << bloc.globLeve));
		bloc.uppeZlen[tb] = bloc.leng[tb] - bloc.deltZlen[tb];
	}
	bloc.SOLVE();
	cap.release();
	std::string log = cap.buf.str();
	bool erro = log.find("(B200): ERROR") != std::string::npos;
	std::cout << std::setprecision(17) << "{\"example\":\"BLOCK\",\"impl\":\"b200\",\"globLeve\":" << glob << ",\"muscSett\":" << musc
		<< ",\"error\":" << (erro ? "true" : "false") << ",\"iterNumbReco\":" << bloc.iterNumbReco << ",\"disp_norm\":[";
	for(long tv = 0; tv < bloc.resuDisp.size(); tv ++) std::cout << (tv ? "," : "") << bloc.resuDisp[tv].norm();
	std::cout << "],\"total_s\":" << now_s() - t0 << "}" << std::endl;
	if(erro){
		size_t p = log.find("(B200): ERROR");
		std::cerr << log.substr(p > 80 ? p - 80 : 0, 400) << std::endl;
	}
	return erro ? 1 : 0;
}
