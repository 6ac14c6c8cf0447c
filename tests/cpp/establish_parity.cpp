// establish_parity.cpp -- TEST INFRASTRUCTURE (CPU, needs /root/reference; built and run by
// tests/test_overlay_compile.py).  SURVEY.md §8 row a2: MGPIS::ESTABLISH (MGPIS.h:40-53) splits every consStif[l] into
// strictly-lower, diagonal and strictly-upper sparse matrices; other reference code copies those members by name
// (MULTIGRID.h:133-138), so the overlay class (ddpca-admm_b200/host/MGPIS.h) has to fill them exactly as the reference
// does.  Both classes are compiled into this one program under different names and compared array by array on
// random hierarchies with ragged rows, explicitly stored zeros and rows without a stored diagonal entry.
// ESTABLISH touches no device: this runs without a GPU.
#include "PREP.h"
#define MGPIS MGPIS_REF
#include <MGPIS.h>                                   // the reference's (-I/root/reference)
#undef MGPIS
#undef _MGPIS_H
#define MGPIS MGPIS_OVL
#include "../../ddpca-admm_b200/host/MGPIS.h"        // the overlay
#undef MGPIS
#include <chrono>
#include <random>

typedef Eigen::SparseMatrix<double, Eigen::RowMajor> SPM_R;

static SPM_R random_level(long n, std::mt19937 &rng) {
	std::uniform_real_distribution<double> val(-1.0, 1.0);
	std::uniform_int_distribution<long> col(0, n - 1), len(0, 9);
	std::vector<Eigen::Triplet<double>> trip;
	for (long i = 0; i < n; i++) {
		if (i % 7 != 3) trip.emplace_back(i, i, 4.0 + val(rng));          // every 7th row: no stored diagonal
		const long k = len(rng);
		for (long t = 0; t < k; t++) {
			const long j = col(rng);
			if (j != i) trip.emplace_back(i, j, (t == 0 && i % 5 == 0) ? 0.0 : val(rng));   // some explicit zeros
		}
	}
	SPM_R A(n, n);
	A.setFromTriplets(trip.begin(), trip.end());
	return A;
}

static bool same(const SPM_R &a0, const SPM_R &b0, const char *what, long level) {
	SPM_R a = a0, b = b0;
	a.makeCompressed();
	b.makeCompressed();
	bool ok = a.rows() == b.rows() && a.cols() == b.cols() && a.nonZeros() == b.nonZeros();
	for (long i = 0; ok && i <= a.rows(); i++) ok = a.outerIndexPtr()[i] == b.outerIndexPtr()[i];
	for (long p = 0; ok && p < a.nonZeros(); p++) ok = a.innerIndexPtr()[p] == b.innerIndexPtr()[p] && a.valuePtr()[p] == b.valuePtr()[p];
	if (!ok) std::cerr << what << "[" << level << "] differs" << std::endl;
	return ok;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// `establish_parity bench [n] [entries per row]`: wall time of both ESTABLISH on one large level (and equality again)
static int bench(long n, long perRow) {
	std::mt19937 rng(7);
	std::uniform_real_distribution<double> val(-1.0, 1.0);
	std::vector<Eigen::Triplet<double>> trip;
	trip.reserve(n * (perRow + 1));
	for (long i = 0; i < n; i++) {
		trip.emplace_back(i, i, 4.0);
		for (long t = 1; t <= perRow / 2; t++) { if (i - 3 * t >= 0) trip.emplace_back(i, i - 3 * t, val(rng)); if (i + 3 * t < n) trip.emplace_back(i, i + 3 * t, val(rng)); }
	}
	MGPIS_REF ref;
	MGPIS_OVL ovl;
	ref.maxiLeve = ovl.maxiLeve = 0;
	ref.consStif.resize(1);
	ref.consStif[0].resize(n, n);
	ref.consStif[0].setFromTriplets(trip.begin(), trip.end());
	ovl.consStif = ref.consStif;
	double t0 = now_s();
	ref.ESTABLISH();
	double t1 = now_s();
	ovl.ESTABLISH();
	double t2 = now_s();
	bool ok = same(ref.consLowe[0], ovl.consLowe[0], "consLowe", 0) && same(ref.consDiag[0], ovl.consDiag[0], "consDiag", 0) && same(ref.consUppe[0], ovl.consUppe[0], "consUppe", 0);
	std::cout << (ok ? "OK" : "MISMATCH") << " rows " << n << " entries " << ref.consStif[0].nonZeros() << ": reference ESTABLISH " << t1 - t0 << " s, overlay " << t2 - t1 << " s" << std::endl;
	return ok ? 0 : 1;
}

int main(int argc, char **argv) {
	if (argc > 1 && std::string(argv[1]) == "bench") return bench(argc > 2 ? std::atol(argv[2]) : 500000, argc > 3 ? std::atol(argv[3]) : 80);
	std::mt19937 rng(12345);
	bool ok = true;
	long compared = 0;
	for (long rep = 0; rep < 4; rep++) {
		const long maxiLeve = rep;                                          // 1 .. 4 levels
		MGPIS_REF ref;
		MGPIS_OVL ovl;
		ref.maxiLeve = ovl.maxiLeve = maxiLeve;
		ref.consStif.resize(maxiLeve + 1);
		ref.realProl.resize(maxiLeve);
		long n = 5 + 3 * rep;
		for (long l = 0; l <= maxiLeve; l++) {
			ref.consStif[l] = random_level(n, rng);
			if (l < maxiLeve) {
				SPM_R P(3 * n + 1, n);
				std::vector<Eigen::Triplet<double>> trip;
				for (long i = 0; i < 3 * n + 1; i++) trip.emplace_back(i, i % n, 1.0 / (1 + i % 3));
				P.setFromTriplets(trip.begin(), trip.end());
				ref.realProl[l] = P;
			}
			n = 3 * n + 1;
		}
		ovl.consStif = ref.consStif;
		ovl.realProl = ref.realProl;
		if (ref.ESTABLISH() != 1 || ovl.ESTABLISH() != 1) { std::cerr << "ESTABLISH returned an error" << std::endl; return 1; }
		ok = ok && (long)ovl.consLowe.size() == maxiLeve + 1 && (long)ovl.consDiag.size() == maxiLeve + 1 && (long)ovl.consUppe.size() == maxiLeve + 1;
		for (long l = 0; ok && l <= maxiLeve; l++) {
			ok = same(ref.consLowe[l], ovl.consLowe[l], "consLowe", l) && same(ref.consDiag[l], ovl.consDiag[l], "consDiag", l)
				&& same(ref.consUppe[l], ovl.consUppe[l], "consUppe", l) && same(ref.consStif[l], ovl.consStif[l], "consStif", l);
			compared++;
		}
		for (long l = 0; ok && l < maxiLeve; l++) ok = same(ref.realProl[l], ovl.realProl[l], "realProl", l);
		// a copy keeps the members (MULTIGRID::COPY, MULTIGRID.h:133-138) and starts without a device hierarchy
		MGPIS_OVL copy = ovl;
		for (long l = 0; ok && l <= maxiLeve; l++) ok = same(copy.consLowe[l], ovl.consLowe[l], "copy.consLowe", l);
	}
	std::cout << (ok ? "OK" : "MISMATCH") << " levels compared: " << compared << std::endl;
	return ok ? 0 : 1;
}
