// ref_capture.h -- helpers shared by the reference-built drivers: capture the
// reference's std::cout chatter, pull "#Iteration: k" out of it (the reference
// returns no iteration count, MGPIS.h:221 prints iterNumb-1), wall-clock timer.
//
// The reference prints from inside its OpenMP loops (MCONTACT.h:2511-2532 -> MGPIS.h:164-221,
// MCONTACT.h:838-847 ...).  On the real std::cout that is serialised by stdio; a std::stringstream
// put in its place is NOT safe for concurrent writers (round-1 drivers did that: racing appends corrupt
// the heap and showed up as sporadic aborts / wrong dumps of the "reference run" on many-core boxes).
// COUT_CAPTURE therefore installs an unbuffered stream buffer whose every write takes a mutex.
#ifndef REF_CAPTURE_H
#define REF_CAPTURE_H
#include <chrono>
#include <iostream>
#include <mutex>
#include <sstream>
#include <streambuf>
#include <string>

class LOCKED_TEXT_BUF : public std::streambuf {
public:
	std::string text() { std::lock_guard<std::mutex> g(mtx_); return text_; }
	std::string str() { return text(); }   // like std::stringstream::str()
	void str(const std::string &t) { std::lock_guard<std::mutex> g(mtx_); text_ = t; }   // reset, like std::stringstream::str("")
protected:
	std::streamsize xsputn(const char *s, std::streamsize n) override {
		std::lock_guard<std::mutex> g(mtx_);
		text_.append(s, (size_t)n);
		return n;
	}
	int overflow(int c) override {
		if (c != EOF) { std::lock_guard<std::mutex> g(mtx_); text_.push_back((char)c); }
		return c;
	}
private:
	std::mutex mtx_;
	std::string text_;
};

struct COUT_CAPTURE {
	LOCKED_TEXT_BUF buf;
	std::streambuf *old;
	COUT_CAPTURE() { old = std::cout.rdbuf(&buf); }
	~COUT_CAPTURE() { release(); }
	void release() { if (old) { std::cout.rdbuf(old); old = nullptr; } }
	// last "#Iteration: k" in the captured text, +1 (== iterNumb of CG_SOLV); -1000 if absent
	long last_iteration_plus1() {
		std::string s = buf.text();
		size_t p = s.rfind("#Iteration: ");
		if (p == std::string::npos) return -1000;
		return std::stol(s.substr(p + 12)) + 1;
	}
};

static inline double now_s() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#endif
