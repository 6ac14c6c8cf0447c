// block_lagrange.cpp -- TEST / ORACLE INFRASTRUCTURE.  One source, two builds (lagrange_tap.h):
// oracle/_ref/block_lagrange (the untouched reference) and host/_bin/block_lagrange_b200 (MGPIS overlay).
// Runs the reference's BLOCK example on its dual-mortar path, menu entry 3 of examples/BLOCK.cpp:96-102:
// SOLVE(2) -> MCONTACT::LAGRANGE(1) (MCONTACT.h:2847-3701): per active-set step the condensed system K is
// assembled on the host (:3413), a multigrid hierarchy is rebuilt for it (:3418-3560), and
// `mgpi.ESTABLISH(); mgpi.BiCGSTAB_SOLV(1, F, U_1)` (:3561-3562) solves it -- the one call that belongs to
// the hot path.  Prints one JSON line: number of active-set steps, BiCGSTAB
// iteration count of every step, norms of the resulting displacements.
//
// The first BiCGSTAB call can be dumped (--out f.ddpk): the condensed hierarchy, its right-hand side and the
// solution become a fixture for the oracle / device parity tests (tests/golden/block_lagrange.ddpk.gz is the
// reference build's).  Both builds time the two MGPIS calls of every active-set step.
//
// usage: block_lagrange --glob G [--divi a,b,c] [--out f.ddpk [--skew s]]   (--skew: lagrange_tap.h, SKEW_VARIANT)
#include "lagrange_tap.h"
#include "MCONTACT.h"
#include "examples/BLOCK.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/BLOCK.cpp:39-40
	omp_set_dynamic(1);
	long glob = 2;
	std::vector<long> divi;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--glob") glob = std::stol(next());
		else if (a == "--divi") { std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) divi.push_back(std::stol(t)); }
		else if (a == "--out") g_lagrOut = next();
		else if (a == "--skew") g_lagrSkew = std::stod(next());
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	BLOCK bloc;
	bloc.domaNumb = {1, 1, 1};            // examples/BLOCK.cpp:96-102 (menu 3)
	if (divi.size() == 3) bloc.diviNumb = divi;
	bloc.globLeve = glob;                 // constructor derives these from globLeve (examples/BLOCK.h:48-53)
	for (long tb = 0; tb < 3; tb++) {
		bloc.deltZlen[tb] = bloc.leng[tb] / (bloc.diviNumb[tb] * (1 << bloc.globLeve));
		bloc.uppeZlen[tb] = bloc.leng[tb] - bloc.deltZlen[tb];
	}
	bloc.SOLVE(2);
	cap.release();
	std::string log = cap.buf.str();
	bool erro = false;
	std::string repo = LAGR_REPORT(log, erro);
	std::cout << "{\"example\":\"BLOCK\",\"globLeve\":" << glob << "," << repo << "," << LAGR_DISP(bloc)
		<< ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	if (erro) {
		size_t p = log.find("ERROR");
		std::cerr << log.substr(p > 200 ? p - 200 : 0, 600) << std::endl;
	}
	return erro ? 1 : 0;
}
