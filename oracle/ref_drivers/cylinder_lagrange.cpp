// cylinder_lagrange.cpp -- TEST / ORACLE INFRASTRUCTURE.  One source, two builds (lagrange_tap.h):
// oracle/_ref/cylinder_lagrange (the untouched reference) and host/_bin/cylinder_lagrange_b200 (MGPIS overlay).
// The reference's Hertzian-contact example on its dual-mortar path, menu entry 3 of examples/CYLINDER.cpp:85-90:
// CYLINDER_1 with copyNumb = 1, SOLVE(2) -> MCONTACT::LAGRANGE(1) (MCONTACT.h:2847-3701).  Unlike the BLOCK patch
// test the contact zone is not known in advance: the active set changes over several steps, each with a newly
// condensed system, a rebuilt hierarchy and one `mgpi.ESTABLISH(); mgpi.BiCGSTAB_SOLV(1, F, U_1)` (:3561-3562).
// The example's sizes are globInho=3, globHomo=0, locaLeve=7 (CYLINDER_1.h:53-55); --inho/--homo/--loca reduce them.
//
// usage: cylinder_lagrange [--inho I] [--homo H] [--loca L] [--out f.ddpk]
#include "lagrange_tap.h"
#include "MCONTACT.h"
#include "examples/CYLINDER_1.h"

int main(int argc, char **argv) {
	omp_set_nested(1);   // examples/CYLINDER.cpp:39-40
	omp_set_dynamic(1);
	long inho = 3, homo = 0, loca = 5;
	for (int i = 1; i < argc; i++) {
		std::string a = argv[i];
		auto next = [&]() { return std::string(argv[++i]); };
		if (a == "--inho") inho = std::stol(next());
		else if (a == "--homo") homo = std::stol(next());
		else if (a == "--loca") loca = std::stol(next());
		else if (a == "--out") g_lagrOut = next();
		else { std::cerr << "unknown arg " << a << std::endl; return 2; }
	}
	double t0 = now_s();
	COUT_CAPTURE cap;
	CYLINDER_1 cyli;
	cyli.copyNumb = 1;                    // examples/CYLINDER.cpp:86-88
	cyli.globInho = inho;
	cyli.globHomo = homo;
	cyli.locaLeve = loca;
	cyli.SOLVE(2);
	cap.release();
	std::string log = cap.buf.str();
	bool erro = false;
	std::string repo = LAGR_REPORT(log, erro);
	std::cout << "{\"example\":\"CYLINDER_1\",\"globInho\":" << inho << ",\"globHomo\":" << homo << ",\"locaLeve\":" << loca << ","
		<< repo << "," << LAGR_DISP(cyli) << ",\"total_s\":" << now_s() - t0 << "}" << std::endl;
	if (erro) {
		size_t p = log.find("ERROR");
		std::cerr << log.substr(p > 200 ? p - 200 : 0, 600) << std::endl;
	}
	return erro ? 1 : 0;
}
