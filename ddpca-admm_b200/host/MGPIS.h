// MGPIS.h -- B200 overlay of the reference's MGPIS.h (class MGPIS, /root/reference/MGPIS.h:8-38).
//
// Drop-in use: compile any reference example unchanged with
//     g++ ... -include <repo>/ddpca-admm_b200/host/MGPIS.h -I<repo>/include -I<reference> \
//         examples/BEAM.cpp -L<repo>/ddpca-admm_b200/lib -lddpca_b200
// The include guard below is the reference's own (_MGPIS_H), so the `#include "MGPIS.h"` inside
// the reference's MULTIGRID.h / MCONTACT.h becomes a no-op and every `MGPIS mgpi` member in the
// reference is this class.  The public surface is kept verbatim (members maxiLeve, realProl,
// consStif, consLowe, consDiag, consUppe are read and written by MULTIGRID.h:102,133-138,
// 1214-1251 and MCONTACT.h:830,853,1539-1542; method signatures as MGPIS.h:16-37); the bodies
// forward to the C ABI of include/ddpca_b200.h.  No CPU fallback: if a device call fails the error is
// printed in the reference's style ("ERROR") and the process ends (MGPIS::DEVICE_FAILURE below).
#ifndef _MGPIS_H
#define _MGPIS_H

#include "PREP.h"
#include "ddpca_b200.h"

#include <cstdlib>
#include <omp.h>
#include <memory>

class MGPIS{
public:
	/*********************************************************************************************/
	long maxiLeve;                                                              // MGPIS.h:12
	std::vector<Eigen::SparseMatrix<double,Eigen::RowMajor>> realProl;          // MGPIS.h:13
	std::vector<Eigen::SparseMatrix<double,Eigen::RowMajor>> consStif;          // MGPIS.h:15
	long ESTABLISH();                                                           // MGPIS.h:16
	long MULT_SOLV(const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu); // MGPIS.h:18
	long CG_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu);       // :22
	long GMRES_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu);    // :24
	long BiCGSTAB_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu); // :26
	/*********************************************************************************************/
	std::vector<Eigen::SparseMatrix<double,Eigen::RowMajor>> consLowe;          // MGPIS.h:29
	std::vector<Eigen::SparseMatrix<double,Eigen::RowMajor>> consDiag;          // MGPIS.h:31
	std::vector<Eigen::SparseMatrix<double,Eigen::RowMajor>> consUppe;          // MGPIS.h:33
	long MULT_VCYC(long tempLeve, const Eigen::VectorXd &righHand,
		Eigen::VectorXd &resuSolu, const DIRE_SOLV &direSolv
	);                                                                          // MGPIS.h:35-37
	/***************************** additions of the B200 build ***********************************/
	MGPIS() : maxiLeve(-1) {}
	// the device hierarchy is owned by one object; copies (MULTIGRID::COPY, vector growth)
	// start without one and build their own on first use
	MGPIS(const MGPIS &o) : maxiLeve(o.maxiLeve), realProl(o.realProl), consStif(o.consStif),
		consLowe(o.consLowe), consDiag(o.consDiag), consUppe(o.consUppe) {}
	MGPIS &operator=(const MGPIS &o){
		if(this != &o){
			maxiLeve = o.maxiLeve; realProl = o.realProl; consStif = o.consStif;
			consLowe = o.consLowe; consDiag = o.consDiag; consUppe = o.consUppe;
			devi.reset();
		}
		return *this;
	}
	// device hierarchy (built lazily from consStif / realProl); nullptr + message on failure
	ddpca_mg *DEVICE_HANDLE();
	// hand the hierarchy over to another owner (ddpca_admm_set_macro_mg takes ownership)
	ddpca_mg *RELEASE_HANDLE(){ ddpca_mg *h = DEVICE_HANDLE(); if(devi) devi->h = nullptr; devi.reset(); return h; }
	// the hierarchy as the plain arrays of the C ABI (Eigen's compressed RowMajor storage of the members);
	// they stay valid as long as consStif / realProl are not modified
	struct POINTERS{
		std::vector<int> n;
		std::vector<const int*> rp, ci, prp, pci;
		std::vector<const double*> va, pva;
	};
	void HIERARCHY_POINTERS(POINTERS &poin);
	static int DEVICE(){ const char *e = std::getenv("DDPCA_DEVICE"); return e ? std::atoi(e) : 0; }
	// DDPCA_SMOOTHER=lex reproduces the reference's lexicographic sweeps exactly (slow),
	// default mc = multicolour ordering of the same symmetric Gauss-Seidel (include/ddpca_b200.h)
	static int SMOOTHER(){
		const char *e = std::getenv("DDPCA_SMOOTHER");
		return (e && (e[0] == 'l' || e[0] == 'L')) ? DDPCA_SMOOTH_LEX : DDPCA_SMOOTH_MC;
	}
	long lastIterNumb = 0;      // iterNumb of the last CG_SOLV / BiCGSTAB_SOLV (the reference only prints it)
	// A device call failed.  The reference's callers ignore return values and would go on with a
	// zero "solution" -- MCONTACT::LAGRANGE's active-set loop (MCONTACT.h:3690-3698) then never ends.  There is no
	// CPU fallback, so the message goes to std::cout in the reference's style and to std::cerr, and the process
	// ends with status 3 unless DDPCA_CONTINUE_ON_ERROR=1 asks for the reference's "print and return -1".
	static long DEVICE_FAILURE(const char *where){
		std::cout << where << " (B200): ERROR " << ddpca_last_error() << std::endl;
		const char *e = std::getenv("DDPCA_CONTINUE_ON_ERROR");
		if(!(e && e[0] == '1')){
			std::cerr << where << " (B200): ERROR " << ddpca_last_error()
				<< " -- no CPU fallback, terminating (DDPCA_CONTINUE_ON_ERROR=1 to return -1 instead)" << std::endl;
			std::cout.flush();
			std::exit(3);
		}
		return -1;
	}
private:
	struct DEVI{
		ddpca_mg *h = nullptr;
		~DEVI(){ if(h) ddpca_mg_destroy(h); }
	};
	std::shared_ptr<DEVI> devi;
};

long MGPIS::ESTABLISH(){
	// the L/D/U members stay available to the reference code that copies them by name
	// (MULTIGRID.h:136-138); the device keeps its own single-copy layout.  Same result as
	// MGPIS.h:40-53 array for array (tests/cpp/establish_parity.cpp) -- strictly-lower and
	// strictly-upper entries of every row in their stored order, one diagonal entry per row
	// (0.0 where none is stored: coeff(tj,tj)) -- but written straight into the compressed
	// arrays in one counting and one filling pass per level, rows in parallel unless the
	// caller is already inside a parallel region (MCONTACT::ESTABLISH loops over bodies,
	// MCONTACT.h:812-825), instead of two triangular views, n binary searches and a triplet sort.
	typedef Eigen::SparseMatrix<double,Eigen::RowMajor> SPMA;
	consLowe.resize(maxiLeve + 1);
	consDiag.resize(maxiLeve + 1);
	consUppe.resize(maxiLeve + 1);
	const bool rowsPara = !omp_in_parallel();
	for(long ti = 0; ti <= maxiLeve; ti ++){
		SPMA &stif = consStif[ti];
		stif.makeCompressed();
		const long rowNumb = stif.rows();
		const int *outeStif = stif.outerIndexPtr();
		const int *inneStif = stif.innerIndexPtr();
		const double *valuStif = stif.valuePtr();
		SPMA lowe(rowNumb, stif.cols()), diag(rowNumb, stif.cols()), uppe(rowNumb, stif.cols());
		//entries below / above the diagonal per row, then prefix sums
		int *outeLowe = lowe.outerIndexPtr(), *outeDiag = diag.outerIndexPtr(), *outeUppe = uppe.outerIndexPtr();
		outeLowe[0] = 0; outeDiag[0] = 0; outeUppe[0] = 0;
		#pragma omp parallel for schedule(static) if(rowsPara)
		for(long tj = 0; tj < rowNumb; tj ++){
			int loweNumb = 0, uppeNumb = 0;
			for(int tk = outeStif[tj]; tk < outeStif[tj + 1]; tk ++){
				loweNumb += (inneStif[tk] < tj);
				uppeNumb += (inneStif[tk] > tj);
			}
			outeLowe[tj + 1] = loweNumb;
			outeUppe[tj + 1] = uppeNumb;
			outeDiag[tj + 1] = 1;
		}
		for(long tj = 0; tj < rowNumb; tj ++){
			outeLowe[tj + 1] += outeLowe[tj];
			outeUppe[tj + 1] += outeUppe[tj];
			outeDiag[tj + 1] += outeDiag[tj];
		}
		lowe.resizeNonZeros(outeLowe[rowNumb]);
		uppe.resizeNonZeros(outeUppe[rowNumb]);
		diag.resizeNonZeros(rowNumb);
		int *inneLowe = lowe.innerIndexPtr(), *inneDiag = diag.innerIndexPtr(), *inneUppe = uppe.innerIndexPtr();
		double *valuLowe = lowe.valuePtr(), *valuDiag = diag.valuePtr(), *valuUppe = uppe.valuePtr();
		#pragma omp parallel for schedule(static) if(rowsPara)
		for(long tj = 0; tj < rowNumb; tj ++){
			int loweCurs = outeLowe[tj], uppeCurs = outeUppe[tj];
			double diagValu = 0.0;
			for(int tk = outeStif[tj]; tk < outeStif[tj + 1]; tk ++){
				if(inneStif[tk] < tj){
					inneLowe[loweCurs] = inneStif[tk];
					valuLowe[loweCurs ++] = valuStif[tk];
				}
				else if(inneStif[tk] > tj){
					inneUppe[uppeCurs] = inneStif[tk];
					valuUppe[uppeCurs ++] = valuStif[tk];
				}
				else{
					diagValu = valuStif[tk];
				}
			}
			inneDiag[tj] = tj;
			valuDiag[tj] = diagValu;
		}
		consLowe[ti].swap(lowe);
		consDiag[ti].swap(diag);
		consUppe[ti].swap(uppe);
	}
	for(long ti = 0; ti < maxiLeve; ti ++){
		realProl[ti].makeCompressed();
	}
	devi.reset();
	return 1;
}

void MGPIS::HIERARCHY_POINTERS(POINTERS &poin){
	const long nlev = maxiLeve + 1;
	poin.n.resize(nlev);
	poin.rp.resize(nlev); poin.ci.resize(nlev); poin.va.resize(nlev);
	poin.prp.resize(nlev); poin.pci.resize(nlev); poin.pva.resize(nlev);
	for(long ti = 0; ti < nlev; ti ++){
		if(!consStif[ti].isCompressed()){
			consStif[ti].makeCompressed();
		}
		poin.n[ti] = consStif[ti].rows();
		poin.rp[ti] = consStif[ti].outerIndexPtr();
		poin.ci[ti] = consStif[ti].innerIndexPtr();
		poin.va[ti] = consStif[ti].valuePtr();
	}
	for(long ti = 0; ti + 1 < nlev; ti ++){
		if(!realProl[ti].isCompressed()){
			realProl[ti].makeCompressed();
		}
		poin.prp[ti] = realProl[ti].outerIndexPtr();
		poin.pci[ti] = realProl[ti].innerIndexPtr();
		poin.pva[ti] = realProl[ti].valuePtr();
	}
}

ddpca_mg *MGPIS::DEVICE_HANDLE(){
	if(devi && devi->h){
		return devi->h;
	}
	const long nlev = maxiLeve + 1;
	POINTERS poin;
	HIERARCHY_POINTERS(poin);
	ddpca_mg *h = nullptr;
	if(ddpca_mg_create(DEVICE(), nlev, poin.n.data(), poin.rp.data(), poin.ci.data(), poin.va.data(),
		poin.prp.data(), poin.pci.data(), poin.pva.data(), SMOOTHER(), &h) != 0){
		return nullptr;   // the caller reports (DEVICE_FAILURE)
	}
	devi = std::make_shared<DEVI>();
	devi->h = h;
	return h;
}

long MGPIS::MULT_VCYC(long tempLeve, const Eigen::VectorXd &righHand,
	Eigen::VectorXd &resuSolu, const DIRE_SOLV &direSolv){
	(void)direSolv;// level 0 is solved with the factorisation held on the device
	ddpca_mg *h = DEVICE_HANDLE();
	if(h == nullptr || ddpca_mg_vcycle(h, tempLeve, righHand.data(), resuSolu.data()) != 0){
		return DEVICE_FAILURE("MGPIS::MULT_VCYC");
	}
	return 1;
}

long MGPIS::CG_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu){
	std::cout << "MGPIS::CG_SOLV";
	if(precSwit == 0){
		std::cout << " (diagonal preconditioner)";
	}
	else if(precSwit == 1){
		std::cout << " (multigrid preconditioner)";
	}
	OUTPUT_TIME("");
	resuSolu = Eigen::VectorXd::Zero(consStif[maxiLeve].rows());
	ddpca_mg *h = DEVICE_HANDLE();
	long iterNumb = 0;
	double resiNorm = 0.0, toleLimi = 0.0;
	if(h == nullptr || ddpca_mg_pcg(h, precSwit, totaForc.data(), resuSolu.data(), 1.0E-14,
		resuSolu.rows(), &iterNumb, &resiNorm, &toleLimi) != 0){
		return DEVICE_FAILURE("MGPIS::CG_SOLV");
	}
	lastIterNumb = iterNumb;
	std::cout << "#Iteration: " << iterNumb - 1
		<< ", residual: " << resiNorm << "/" << toleLimi;
	OUTPUT_TIME(":");
	return 1;
}

long MGPIS::BiCGSTAB_SOLV(long precSwit,
	const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu){
	std::cout << "MGPIS::BiCGSTAB_SOLV";
	if(precSwit == 0){
		std::cout << " (diagonal preconditioner)";
	}
	else if(precSwit == 1){
		std::cout << " (multigrid preconditioner)";
	}
	OUTPUT_TIME("");
	resuSolu = Eigen::VectorXd::Zero(consStif[maxiLeve].rows());
	ddpca_mg *h = DEVICE_HANDLE();
	long iterNumb = 0;
	double resiNorm = 0.0, toleLimi = 0.0;
	if(h == nullptr || ddpca_mg_bicgstab(h, precSwit, totaForc.data(), resuSolu.data(), 1.0E-14,
		resuSolu.rows(), &iterNumb, &resiNorm, &toleLimi) != 0){
		return DEVICE_FAILURE("MGPIS::BiCGSTAB_SOLV");
	}
	lastIterNumb = iterNumb;
	std::cout << "#Iteration: " << iterNumb - 1
		<< ", residual: " << resiNorm << "/" << toleLimi;
	OUTPUT_TIME(":");
	return 1;
}

long MGPIS::MULT_SOLV(const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu){
	OUTPUT_TIME("MGPIS::MULT_SOLV");
	resuSolu = Eigen::VectorXd::Zero(consStif[maxiLeve].rows());
	ddpca_mg *h = DEVICE_HANDLE();
	long iterNumb = 0;
	double resiNorm = 0.0;
	if(h == nullptr || ddpca_mg_mult_solv(h, totaForc.data(), resuSolu.data(), &iterNumb, &resiNorm) != 0){
		return DEVICE_FAILURE("MGPIS::MULT_SOLV");
	}
	std::cout << "#Iteration: " << iterNumb << ", residual: "
		<< resiNorm << "/" << 1.0E-14 * totaForc.norm();
	OUTPUT_TIME(":");
	return 1;
}

long MGPIS::GMRES_SOLV(long precSwit, const Eigen::VectorXd &totaForc, Eigen::VectorXd &resuSolu){
	std::cout << "MGPIS::GMRES_SOLV";
	if(precSwit == 0){
		std::cout << " (diagonal preconditioner)";
	}
	else if(precSwit == 1){
		std::cout << " (multigrid preconditioner)";
	}
	OUTPUT_TIME("");
	resuSolu = Eigen::VectorXd::Zero(consStif[maxiLeve].rows());
	ddpca_mg *h = DEVICE_HANDLE();
	long iterNumb = 0;
	double resiNorm = 0.0, toleLimi = 0.0;
	if(h == nullptr || ddpca_mg_gmres(h, precSwit, totaForc.data(), resuSolu.data(),
		&iterNumb, &resiNorm, &toleLimi) != 0){
		return DEVICE_FAILURE("MGPIS::GMRES_SOLV");
	}
	lastIterNumb = iterNumb;
	std::cout << "#Iteration: " << iterNumb << ", residual: " << resiNorm << "/" << toleLimi;
	OUTPUT_TIME(":");
	return 1;
}

#endif
