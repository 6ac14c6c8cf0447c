// ref_capture.h -- helpers shared by the reference-built drivers: capture the
// reference's std::cout chatter, pull "#Iteration: k" out of it (the reference
// returns no iteration count, MGPIS.h:221 prints iterNumb-1), wall-clock timer.
#ifndef REF_CAPTURE_H
#define REF_CAPTURE_H
#include <chrono>
#include <iostream>
#include <sstream>
#include <string>

struct COUT_CAPTURE {
	std::stringstream buf;
	std::streambuf *old;
	COUT_CAPTURE() { old = std::cout.rdbuf(buf.rdbuf()); }
	~COUT_CAPTURE() { release(); }
	void release() { if (old) { std::cout.rdbuf(old); old = nullptr; } }
	// last "#Iteration: k" in the captured text, +1 (== iterNumb of CG_SOLV); -1000 if absent
	long last_iteration_plus1() const {
		std::string s = buf.str();
		size_t p = s.rfind("#Iteration: ");
		if (p == std::string::npos) return -1000;
		return std::stol(s.substr(p + 12)) + 1;
	}
};

static inline double now_s() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#endif
