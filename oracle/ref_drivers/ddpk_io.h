// ddpk_io.h -- tiny named-array container ("DDPK") used to move operators and
// vectors between the reference-built drivers (oracle/_ref/*), the C oracle,
// the python tests and bench.py.  TEST / ORACLE INFRASTRUCTURE ONLY.
//
// Layout: 8-byte magic "DDPK0001", then records
//   u32 name_len | name bytes | u32 dtype (0=f64,1=i32,2=i64) | u64 count |
//   zero padding to an 8-byte file offset | raw little-endian data
// Sparse matrices are stored as four records: <name>.shape (i64[2]),
// <name>.rowptr (i32[rows+1]), <name>.colidx (i32[nnz]), <name>.val (f64[nnz])
// which is exactly Eigen's compressed RowMajor storage.
#ifndef DDPK_IO_H
#define DDPK_IO_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

class DDPK_WRITER {
public:
	explicit DDPK_WRITER(const std::string &path) {
		fp = std::fopen(path.c_str(), "wb");
		if (!fp) { std::perror(path.c_str()); std::exit(2); }
		std::fwrite("DDPK0001", 1, 8, fp);
		off = 8;
	}
	~DDPK_WRITER() { if (fp) std::fclose(fp); }
	void raw(const std::string &name, uint32_t dtype, uint64_t count, const void *data, size_t elsz) {
		uint32_t nl = (uint32_t)name.size();
		put(&nl, 4); put(name.data(), nl); put(&dtype, 4); put(&count, 8);
		static const char zeros[8] = {0};
		size_t pad = (8 - off % 8) % 8;
		put(zeros, pad);
		put(data, count * elsz);
	}
	void f64(const std::string &name, const double *p, uint64_t n) { raw(name, 0, n, p, 8); }
	void i32(const std::string &name, const int *p, uint64_t n) { raw(name, 1, n, p, 4); }
	void i64(const std::string &name, const long *p, uint64_t n) { raw(name, 2, n, p, 8); }
	void scalar_i64(const std::string &name, long v) { i64(name, &v, 1); }
	void scalar_f64(const std::string &name, double v) { f64(name, &v, 1); }
#ifdef EIGEN_SPARSEMATRIX_H
	void csr(const std::string &name, const Eigen::SparseMatrix<double, Eigen::RowMajor> &A0) {
		Eigen::SparseMatrix<double, Eigen::RowMajor> A = A0;
		A.makeCompressed();
		long shape[2] = {(long)A.rows(), (long)A.cols()};
		i64(name + ".shape", shape, 2);
		i32(name + ".rowptr", A.outerIndexPtr(), A.rows() + 1);
		i32(name + ".colidx", A.innerIndexPtr(), A.nonZeros());
		f64(name + ".val", A.valuePtr(), A.nonZeros());
	}
	void vec(const std::string &name, const Eigen::VectorXd &v) { f64(name, v.data(), v.size()); }
#endif
private:
	void put(const void *p, size_t n) { if (n) { std::fwrite(p, 1, n, fp); off += n; } }
	FILE *fp;
	size_t off;
};
#endif
