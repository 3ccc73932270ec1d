// boost::iostreams::filtering_{i,o}stream + file_source/file_sink over std::filebuf (no compression filters:
// the reference only pushes them under HAVE_LIBZ, which this build leaves undefined)
#ifndef HU_SHIM_BOOST_IOSTREAMS
#define HU_SHIM_BOOST_IOSTREAMS
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
namespace boost { namespace iostreams {
class file_base {
public:
	file_base(const std::string& fn, std::ios_base::openmode m) : buf_(new std::filebuf()) { buf_->open(fn.c_str(), m); }
	bool is_open() const { return buf_->is_open(); }
	void close() { buf_->close(); }
	std::shared_ptr<std::filebuf> buf_;
};
class file_source : public file_base {
public:
	file_source(const std::string& fn, std::ios_base::openmode m = std::ios_base::in) : file_base(fn, m | std::ios_base::in) {}
};
class file_sink : public file_base {
public:
	file_sink(const std::string& fn, std::ios_base::openmode m = std::ios_base::out) : file_base(fn, m | std::ios_base::out) {}
};
class filtering_istream : public std::istream {
public:
	filtering_istream() : std::istream(0) {}
	void push(const file_source& s) { keep_ = s.buf_; rdbuf(keep_.get()); clear(); }
	void push(std::istream& is) { rdbuf(is.rdbuf()); clear(); }
	bool empty() const { return rdbuf() == 0; }
	bool is_complete() const { return rdbuf() != 0; }
	void pop() { rdbuf(0); keep_.reset(); }
	void reset() { pop(); }
private:
	std::shared_ptr<std::filebuf> keep_;
};
class filtering_ostream : public std::ostream {
public:
	filtering_ostream() : std::ostream(0) {}
	~filtering_ostream() { if (rdbuf()) flush(); }
	void push(const file_sink& s) { keep_ = s.buf_; rdbuf(keep_.get()); clear(); }
	void push(std::ostream& os) { rdbuf(os.rdbuf()); clear(); }
	bool empty() const { return rdbuf() == 0; }
	bool is_complete() const { return rdbuf() != 0; }
	void pop() { if (rdbuf()) flush(); rdbuf(0); keep_.reset(); }
	void reset() { pop(); }
private:
	std::shared_ptr<std::filebuf> keep_;
};
} }
#endif
