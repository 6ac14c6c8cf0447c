// MCONTACT_B200.h -- B200 drop-in for MCONTACT::CONTACT_ANALYSIS (/root/reference/MCONTACT.h:2493-2723).
//
// Include AFTER the reference's MCONTACT.h (and with the MGPIS overlay force-included, see
// INTEGRATION.md).  DDPCA_CONTACT_ANALYSIS(mc) does what mc.CONTACT_ANALYSIS() does -- same state
// members on return (resuDisp, inteAuxi, inteLagr, iterNumbReco), same resuMoni.txt rows, same
// console lines per iteration, same final resuCont_<ts>.txt -- with the loop body on the GPU:
//   1. upload, once, everything MCONTACT::ESTABLISH built (MCONTACT.h:29-46, :864-872) through the
//      C ABI of include/ddpca_b200.h; the host factorisations of ESTABLISH become device solvers;
//   2. iterate ddpca_admm_step(); MONITOR (MCONTACT.h:2725-2845) is restated below on the sums the
//      device returns (the reference's MONITOR needs the full state vectors on the host);
//   3. read the state back.
// Defining DDPCA_HOOK_CONTACT_ANALYSIS before including this file additionally redirects every later
// call `CONTACT_ANALYSIS()` written inside a member function (e.g. examples/BLOCK.h:707) to the
// drop-in, so that the reference's examples run unchanged.
#ifndef _MCONTACT_B200_H
#define _MCONTACT_B200_H

#include "ddpca_b200.h"

#ifndef _MCONTACT_H
#error "include the reference's MCONTACT.h before MCONTACT_B200.h"
#endif

namespace ddpca_host {

typedef Eigen::SparseMatrix<double,Eigen::RowMajor> SPM;

// A device call failed: there is no CPU fallback, and the reference's examples ignore the return value of
// CONTACT_ANALYSIS() and would go on to write result files of an all-zero state.  Same policy as the MGPIS overlay
// (MGPIS::DEVICE_FAILURE): the reference-style ERROR line on std::cout, the same on std::cerr, exit status 3 --
// unless DDPCA_CONTINUE_ON_ERROR=1 asks for the reference's "print and return -1".
inline bool FAIL(const char *what){
	std::cout << "MCONTACT::CONTACT_ANALYSIS (B200): ERROR " << what << ": " << ddpca_last_error() << std::endl;
	const char *e = std::getenv("DDPCA_CONTINUE_ON_ERROR");
	if(!(e && e[0] == '1')){
		std::cerr << "MCONTACT::CONTACT_ANALYSIS (B200): ERROR " << what << ": " << ddpca_last_error()
			<< " -- no CPU fallback, terminating (DDPCA_CONTINUE_ON_ERROR=1 to return -1 instead)" << std::endl;
		std::cout.flush();
		std::exit(3);
	}
	return false;
}

// a factorised DIRE_SOLV (Eigen::SimplicialLDLT) -> device solver; small SPD operators are
// inverted densely on the device instead (no factor needed)
// denseMaxi: interface mass matrices 8192; the coarse problems, solved on every device in every iteration and the
// part of an iteration that does not shrink with more devices, 12288 (one product instead of staged sparse sweeps;
// beyond, the 2 n^3 flops of the inversion cost more at set-up than the product saves over a few dozen iterations)
inline ddpca_ldlt *UPLOAD_SOLVER(int devi, const DIRE_SOLV &solv, const SPM &matr, long denseMaxi = 8192){
	ddpca_ldlt *resu = nullptr;
	if(matr.rows() <= denseMaxi){
		SPM tempMatr = matr;
		tempMatr.makeCompressed();
		if(ddpca_ldlt_create_dense(devi, tempMatr.rows(), tempMatr.outerIndexPtr(),
			tempMatr.innerIndexPtr(), tempMatr.valuePtr(), &resu) != 0){
			return nullptr;
		}
		return resu;
	}
	SPM lowe = solv.matrixL().nestedExpression();// strictly lower, unit diagonal implied
	lowe.makeCompressed();
	Eigen::VectorXd diag = solv.vectorD();
	std::vector<int> perm(solv.rows());
	for(long ti = 0; ti < solv.rows(); ti ++){
		perm[ti] = solv.permutationP().indices()(ti);
	}
	if(ddpca_ldlt_create(devi, lowe.rows(), perm.data(), lowe.outerIndexPtr(),
		lowe.innerIndexPtr(), lowe.valuePtr(), diag.data(), &resu) != 0){
		return nullptr;
	}
	return resu;
}

inline bool UPLOAD_OP(ddpca_admm *hand, long ts, long tv, int opid, const SPM &matr){
	SPM tempMatr = matr;
	tempMatr.makeCompressed();
	return ddpca_admm_set_side_op(hand, ts, tv, opid, tempMatr.rows(), tempMatr.cols(),
		tempMatr.outerIndexPtr(), tempMatr.innerIndexPtr(), tempMatr.valuePtr()) == 0;
}

}// namespace ddpca_host

// devices of the run: DDPCA_DEVICES="0,1,2,3" (one process, several GPUs: the bodies are bin-packed over them by the
// size of their finest operator, include/ddpca_b200.h "one process, several GPUs"), else the single DDPCA_DEVICE
inline std::vector<int> DDPCA_DEVICE_LIST(){
	std::vector<int> resu;
	const char *envi = std::getenv("DDPCA_DEVICES");
	if(envi != nullptr && envi[0] != 0){
		std::stringstream tempStre(envi);
		std::string item;
		while(std::getline(tempStre, item, ',')){
			if(!item.empty()){
				resu.push_back(std::atoi(item.c_str()));
			}
		}
	}
	if(resu.empty()){
		resu.push_back(MGPIS::DEVICE());
	}
	return resu;
}

inline long DDPCA_CONTACT_ANALYSIS(MCONTACT &mc){
	using namespace ddpca_host;
	const std::vector<int> deviList = DDPCA_DEVICE_LIST();
	const long bodyNumb = mc.multGrid.size(), inteNumb = mc.searCont.size();
	const bool macrSwit = ((mc.muscSett >> 0) % 2 == 1);
	const bool elimSwit = ((mc.muscSett >> 1) % 2 == 1);// interface-eliminated coarse problem, MCONTACT.h:2575-2607
	//bodies -> devices (balanced groups; the reference's omp loop over bodies, MCONTACT.h:2511)
	std::vector<double> bodyWeig(bodyNumb);
	std::vector<int> contPair(2 * inteNumb), bodyRank(bodyNumb, 0);
	for(long tv = 0; tv < bodyNumb; tv ++){
		bodyWeig[tv] = mc.multGrid[tv].mgpi.consStif[mc.multGrid[tv].mgpi.maxiLeve].nonZeros();
	}
	for(long ts = 0; ts < inteNumb; ts ++){
		contPair[2 * ts + 0] = mc.contBody[ts][0];
		contPair[2 * ts + 1] = mc.contBody[ts][1];
	}
	ddpca_admm_group *grou = nullptr;
	if(ddpca_partition_bodies(bodyNumb, bodyWeig.data(), inteNumb, contPair.data(), deviList.size(), bodyRank.data()) != 0
		|| ddpca_admm_group_create(deviList.size(), deviList.data(), bodyNumb, inteNumb,
			(macrSwit ? 1 : 0) | (elimSwit ? 2 : 0), bodyRank.data(), &grou) != 0){
		FAIL("create"); return -1;
	}
	const long membNumb = ddpca_admm_group_size(grou);
	bool allGood = true;
	for(long tk = 0; tk < membNumb; tk ++){
		allGood = allGood && ddpca_admm_set_smoother(ddpca_admm_group_member(grou, tk), MGPIS::SMOOTHER()) == 0;
	}
	//******************************** upload (once) ********************************************
	for(long tv = 0; tv < bodyNumb && allGood; tv ++){
		ddpca_admm *hand = ddpca_admm_group_member(grou, bodyRank[tv]);
		MULTIGRID &mugr = mc.multGrid[tv];
		const long maxiLeve = mugr.mgpi.maxiLeve;
		// ADDITIONAL_FORCE as one operator (MULTIGRID.h:1257-1261); OUTP_SUB1 = its transpose + constant
		SPM forcOper = mugr.consOper[maxiLeve] * SPM(mugr.prolOper[maxiLeve].transpose())
			* SPM(mugr.earlTran.transpose());
		forcOper.makeCompressed();
		Eigen::VectorXd dispCons;
		mugr.OUTP_SUB1(Eigen::VectorXd::Zero(forcOper.rows()), dispCons);
		// the hierarchy goes over as plain arrays: ddpca_admm_finalize builds ONE batched device hierarchy for
		// all bodies of a device with the same level count (the members of mgpi stay untouched until then)
		MGPIS::POINTERS poin;
		mugr.mgpi.HIERARCHY_POINTERS(poin);
		if(ddpca_admm_set_body(hand, tv, maxiLeve + 1, poin.n.data(), poin.rp.data(), poin.ci.data(),
			poin.va.data(), poin.prp.data(), poin.pci.data(), poin.pva.data(), 3 * mugr.nodeCoor.size(),
			mugr.consForc.data(), forcOper.outerIndexPtr(), forcOper.innerIndexPtr(),
			forcOper.valuePtr(), dispCons.data()) != 0){
			allGood = FAIL("set_body");
			break;
		}
		if(macrSwit || elimSwit){
			SPM accu = mc.accuProl[tv];
			accu.makeCompressed();
			if(ddpca_admm_set_body_accuprol(hand, tv, accu.rows(), accu.cols(),
				accu.outerIndexPtr(), accu.innerIndexPtr(), accu.valuePtr()) != 0){
				allGood = FAIL("set_body_accuprol");
			}
		}
		if(allGood && elimSwit){
			SPM tran = mc.globTran_D_1[tv];// MCONTACT.h:2583
			tran.makeCompressed();
			if(ddpca_admm_set_body_globtran_d1(hand, tv, tran.rows(), tran.cols(),
				tran.outerIndexPtr(), tran.innerIndexPtr(), tran.valuePtr()) != 0){
				allGood = FAIL("set_body_globtran_d1");
			}
		}
	}
	for(long ts = 0; ts < inteNumb && allGood; ts ++){
		Eigen::VectorXd gapTerm = mc.pemaInpo[ts] * mc.inpoNgap[ts];// MCONTACT.h:2636
		for(long tk = 0; tk < membNumb && allGood; tk ++){// every member knows every interface
			if(ddpca_admm_set_interface(ddpca_admm_group_member(grou, tk), ts, mc.contBody[ts][0], mc.contBody[ts][1],
				mc.fricCoef[ts], mc.searCont[ts].intePoin.size(), gapTerm.data()) != 0){
				allGood = FAIL("set_interface");
			}
		}
		for(long tv = 0; tv < 2 && allGood; tv ++){
			const long ownr = bodyRank[mc.contBody[ts][tv]];// the side lives with its body
			ddpca_admm *hand = ddpca_admm_group_member(grou, ownr);
			const int devi = ddpca_admm_group_device(grou, ownr);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_SYSTTRAN, mc.systTran[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_SYSTTRAN_PENA, mc.systTran_pena[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_INTEMASS, mc.inteMass[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_INTEMASS_PENA, mc.inteMass_pena[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_INPOLAGR, mc.inpoLagr[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_INTEINPO, mc.inteInpo[ts][tv]);
			allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_PEMAINPO_R, mc.pemaInpo_r[ts][tv]);
			if(macrSwit){
				allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_GLOBTRAN, mc.globTran[ts][tv]);
				allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_GLOBTRAN_PENA, mc.globTran_pena[ts][tv]);
				allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_GLOBTRAN_D, mc.globTran_D[ts][tv]);
			}
			if(elimSwit){
				allGood = allGood && UPLOAD_OP(hand, ts, tv, DDPCA_OP_GLOBTRAN_1, mc.globTran_1[ts][tv]);// :2579
			}
			if(!allGood){ FAIL("set_side_op"); break; }
			if(mc.inteMass[ts][tv].rows() >= DIRE_MAXI){// MCONTACT.h:2676-2683, :2698-2703: per-iteration Eigen CG with its
				//diagonal preconditioner, no factor exists -> batched Jacobi-PCG on the device
				if(ddpca_admm_set_side_iterative(hand, ts, tv) != 0){
					allGood = FAIL("set_side_iterative");
				}
				continue;
			}
			ddpca_ldlt *soMa = UPLOAD_SOLVER(devi, mc.inteDiso[ts][tv], mc.inteMass[ts][tv]);
			ddpca_ldlt *soPe = UPLOAD_SOLVER(devi, mc.inteDiso_pena[ts][tv], mc.inteMass_pena[ts][tv]);
			if(soMa == nullptr || soPe == nullptr
				|| ddpca_admm_set_side_solver(hand, ts, tv, DDPCA_SOLVER_MASS, soMa) != 0
				|| ddpca_admm_set_side_solver(hand, ts, tv, DDPCA_SOLVER_MASS_PENA, soPe) != 0){
				allGood = FAIL("set_side_solver");
			}
		}
	}
	for(long tk = 0; tk < membNumb && allGood; tk ++){// the coarse problems are solved redundantly on every device
		ddpca_admm *hand = ddpca_admm_group_member(grou, tk);
		const int devi = ddpca_admm_group_device(grou, tk);
		if(macrSwit){
			if(mc.globCoup.rows() < DIRE_MAXI){// MCONTACT.h:2553-2555
				ddpca_ldlt *soCo = UPLOAD_SOLVER(devi, mc.coarSolv_D, mc.globCoup, 12288);
				if(soCo == nullptr || ddpca_admm_set_macro(hand, mc.globCoup.rows(), mc.baseReco.data(), soCo) != 0){
					allGood = FAIL("set_macro");
				}
			}
			else{// MCONTACT.h:2560-2562: mgpi.CG_SOLV(1, globForc, globSolu), hierarchy of DOUBLE_M (:1538-1670)
				//(COGR_MAXI < DIRE_MAXI, PREP.h:69,73: the Eigen-CG branch :2556-2558 is unreachable)
				MGPIS::POINTERS poin;
				mc.mgpi.HIERARCHY_POINTERS(poin);
				ddpca_mg *mgHand = nullptr;
				if(ddpca_mg_create(devi, mc.mgpi.maxiLeve + 1, poin.n.data(), poin.rp.data(), poin.ci.data(),
					poin.va.data(), poin.prp.data(), poin.pci.data(), poin.pva.data(), MGPIS::SMOOTHER(), &mgHand) != 0
					|| ddpca_admm_set_macro_mg(hand, mc.globCoup.rows(), mc.baseReco.data(), mgHand) != 0){
					allGood = FAIL("set_macro_mg");
				}
			}
		}
		if(allGood && elimSwit){// MCONTACT.h:2576,2588
			if(mc.globCoup_1.rows() < DIRE_MAXI){
				ddpca_ldlt *soCo = UPLOAD_SOLVER(devi, mc.coarSolv_D_1, mc.globCoup_1, 12288);
				if(soCo == nullptr || ddpca_admm_set_macro1(hand, mc.globCoup_1.rows(), mc.baseReco.data(),
					mc.globForc_1.data(), soCo) != 0){
					allGood = FAIL("set_macro1");
				}
			}
			else{// MCONTACT.h:2593-2595: mgpi_1.CG_SOLV(1, globForc, globSolu) (COGR_MAXI < DIRE_MAXI: :2590-2592 unreachable)
				MGPIS::POINTERS poin;
				mc.mgpi_1.HIERARCHY_POINTERS(poin);
				ddpca_mg *mgHand = nullptr;
				if(ddpca_mg_create(devi, mc.mgpi_1.maxiLeve + 1, poin.n.data(), poin.rp.data(), poin.ci.data(),
					poin.va.data(), poin.prp.data(), poin.pci.data(), poin.pva.data(), MGPIS::SMOOTHER(), &mgHand) != 0
					|| ddpca_admm_set_macro1_mg(hand, mc.globCoup_1.rows(), mc.baseReco.data(),
						mc.globForc_1.data(), mgHand) != 0){
					allGood = FAIL("set_macro1_mg");
				}
			}
		}
	}
	if(allGood && ddpca_admm_group_finalize(grou) != 0){
		allGood = FAIL("finalize");
	}
	if(!allGood){
		ddpca_admm_group_destroy(grou);
		return -1;
	}
	//******************************** the loop, MCONTACT.h:2494-2712 ***************************
	const long moniCycl = 10;
	VECTOR2D moniReco(bodyNumb + 4 * inteNumb);
	for(long ti = 0; ti < moniReco.size(); ti ++){
		moniReco[ti].assign(moniCycl, 0.0);
	}
	long tc;
	const long maxiIter = 3000;
	std::ofstream tempOfst(DIRECTORY("resuMoni.txt"), std::ios::out);
	tempOfst << std::setiosflags(std::ios::scientific) << std::setprecision(20);
	std::vector<double> moniRow(ddpca_admm_row_length(ddpca_admm_group_member(grou, 0)));
	for(tc = 0; tc < maxiIter; tc ++){
		std::cout << "The " << tc << "-th iteration";
		OUTPUT_TIME("");
		const int applMacr = ((macrSwit || elimSwit) && tc <= MULT_MAXI) ? 1 : 0;// :2540, :2575
		long cgitNumb = 0;
		if(ddpca_admm_group_step(grou, applMacr, moniRow.data(), &cgitNumb, nullptr) != 0){
			FAIL("step");
			ddpca_admm_group_destroy(grou);
			return -1;
		}
		//stopping criterion: MONITOR (:2725-2845) on the sums computed by the device
		bool tempFlag_0 = (tc >= moniCycl) ? true : false;
		bool tempFlag_1 = true;
		const double critRati_0 = 0.1, critRati_1 = 1.0E-12;
		long tempColu = 0;
		for(long tv = 0; tv < bodyNumb; tv ++){
			const double dispVari = moniRow[tempColu], dispAllo = moniRow[tempColu + 1];
			tempColu += 2;
			moniReco[tv][tc % moniCycl] = dispVari;
			tempOfst << std::setw(30) << dispVari << std::setw(30) << dispAllo;
			if(tc >= moniCycl){
				double dispMedi, dispOsci;
				VECT_MEDI_OSCI(moniReco[tv], dispMedi, dispOsci);
				if(dispOsci > critRati_0 * dispMedi){
					tempFlag_0 = false;
				}
			}
			if(dispVari > critRati_1 * dispAllo){
				tempFlag_1 = false;
			}
		}
		for(long ts = 0; ts < inteNumb; ts ++){
			for(long tv = 0; tv < 2; tv ++){
				const long tempIndi = bodyNumb + 4 * ts + 2 * tv;
				const double auxiVari = moniRow[tempColu], auxiAllo = moniRow[tempColu + 1];
				const double lagrVari = moniRow[tempColu + 2], lagrAllo = moniRow[tempColu + 3];
				tempColu += 4;
				moniReco[tempIndi][tc % moniCycl] = auxiVari;
				moniReco[tempIndi + 1][tc % moniCycl] = lagrVari;
				tempOfst << std::setw(30) << auxiVari << std::setw(30) << auxiAllo
					<< std::setw(30) << lagrVari << std::setw(30) << lagrAllo;
				if(tc >= moniCycl){
					double auxiMedi, auxiOsci;
					VECT_MEDI_OSCI(moniReco[tempIndi], auxiMedi, auxiOsci);
					if(auxiOsci > critRati_0 * auxiMedi){
						tempFlag_0 = false;
					}
				}
				if(auxiVari > critRati_1 * auxiAllo){
					tempFlag_1 = false;
				}
				//the multiplier criteria are evaluated but disabled in the reference (:2822,:2830)
			}
		}
		const double convValu = moniRow[tempColu], convCrit = moniRow[tempColu + 1];
		std::cout << "Cvalu = " << convValu << ", Ccrit = " << convCrit << std::endl;
		tempOfst << std::setw(30) << convValu << std::setw(30) << convCrit;
		tempOfst << std::endl;
		if(tempFlag_0 == true){
			MULT_MAXI = tc;// :2838-2840
		}
		if(tempFlag_1 == true){
			break;// :2709-2711
		}
	}
	tempOfst.close();
	mc.iterNumbReco = tc;
	//******************************** state back to the host members *****************************
	for(long tv = 0; tv < bodyNumb; tv ++){
		mc.resuDisp[tv].resize(3 * mc.multGrid[tv].nodeCoor.size());
		ddpca_admm_get_disp(ddpca_admm_group_member(grou, bodyRank[tv]), tv, mc.resuDisp[tv].data());
	}
	for(long ts = 0; ts < inteNumb; ts ++){
		for(long tv = 0; tv < 2; tv ++){
			ddpca_admm_get_side(ddpca_admm_group_member(grou, bodyRank[mc.contBody[ts][tv]]), ts, tv,
				mc.inteAuxi[ts][tv].data(), mc.inteLagr[ts][tv].data());
		}
		//final contact pressure/traction file, what OUTPUT_PRTR writes each iteration (:2669)
		const long gammSize = ((mc.fricCoef[ts] == 0.0) ? 1 : 3) * mc.searCont[ts].intePoin.size();
		Eigen::VectorXd inpoGamm(gammSize);
		Eigen::VectorXi fricStat(gammSize);
		ddpca_admm_get_gamma(ddpca_admm_group_member(grou, bodyRank[mc.contBody[ts][0]]), ts, inpoGamm.data(), fricStat.data());
		mc.OUTPUT_PRTR(inpoGamm, fricStat, ts);
	}
	ddpca_admm_group_destroy(grou);
	if(tc >= maxiIter){
		std::cout << "Nonconvergence in MCONTACT::CONTACT_ANALYSIS";
	}
	else{
		std::cout << "Converge after " << tc << "-th iterations";
	}
	OUTPUT_TIME("");
	return 1;
}

#ifdef DDPCA_HOOK_CONTACT_ANALYSIS
#define CONTACT_ANALYSIS() DDPCA_CONTACT_ANALYSIS(*this)
#endif

#endif
