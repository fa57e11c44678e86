// manta_module.cpp -- the host layer: a mantaflow-style Python module `manta` whose classes and
// plugin functions have the names, argument names, defaults and error behaviour that
// scenes/flof.py and scenes/ofHelpers.py of the reference (thunil/ofblend) use, with every grid
// resident on the B200 and every operator forwarded to the C ABI (include/flof_b200.h).
//
// ref: the reference generates this surface with its `prep` preprocessor from PYTHON() annotations
// (preprocessor/codegen_python.cpp:27-53, pwrapper/registry.cpp:546-676); here pybind11 plays that
// role.  Mirrored: FluidSolver "Solver" (fluidsolver.h:27-68), Grid4d<T> (grid4d.h:117-271),
// Grid<T>/LevelsetGrid subset (grid.h, levelset.cpp:114-118), vec3/vec4 (pwrapper/pvec3.cpp),
// python/defines.py, and the plugin functions of plugin/optflow4d.cpp, grid4d.cpp, test.cpp, fileio.cpp.
//
// Conventions kept from the reference: arguments by position or by the C++ parameter name; `int`
// parameters accept integral floats (pwrapper/pconvert.cpp:120-132); Vec4 parameters accept a vec4 or
// a 4-sequence (:182-193); every call accepts notiming= (added by the Python shim at the bottom);
// errors surface as RuntimeError (pwrapper/pclass.cpp:50-54).  There is no CPU fallback: the module
// fails to create a Solver without a CUDA device.
#include <unistd.h>
#include <pybind11/pybind11.h>
#include <pybind11/eval.h>
#include <pybind11/stl.h>
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/flof_b200.h"

namespace py = pybind11;

// ------------------------------------------------------------------ context / errors ------
static flof_ctx *g_ctx = nullptr;
static int g_debug_level = 1;

static flof_ctx *ctx()
{
	if (!g_ctx) {
		int dev = 0;
		if (const char *e = getenv("FLOF_DEVICE")) dev = atoi(e);
		else if (const char *e2 = getenv("LOCAL_RANK")) dev = atoi(e2);
		if (flof_ctx_create(&g_ctx, dev) != FLOF_OK)
			throw std::runtime_error(std::string("manta (B200): ") + flof_last_error(nullptr));
	}
	return g_ctx;
}
static void chk(int rc, const char *fn)
{
	if (rc != FLOF_OK) throw std::runtime_error(std::string("Error in ") + fn + ": " + flof_last_error(g_ctx));
}
#define CK(call, fn) chk((call), fn)
static void errMsg(const std::string &m) { throw std::runtime_error(m); }
#define debMsg(level, ...)                      \
	do {                                        \
		if (g_debug_level >= (level)) {         \
			std::ostringstream s__;             \
			s__ << __VA_ARGS__;                 \
			py::print(s__.str());               \
		}                                       \
	} while (0)

// ------------------------------------------------------------------ value types ----------
struct PInt { int v; };  // int parameter that also accepts integral floats (pconvert.cpp:120-132)
namespace pybind11 { namespace detail {
template <> struct type_caster<PInt> {
	PYBIND11_TYPE_CASTER(PInt, const_name("int"));
	bool load(handle src, bool)
	{
		if (!src) return false;
		if (PyBool_Check(src.ptr())) { value.v = src.ptr() == Py_True; return true; }
		if (PyLong_Check(src.ptr())) { value.v = (int)PyLong_AsLong(src.ptr()); return !PyErr_Occurred(); }
		if (PyFloat_Check(src.ptr())) {
			const double a = PyFloat_AsDouble(src.ptr());
			if (std::fabs(a - std::floor(a + 0.5)) > 1e-5) throw std::runtime_error("argument is not an int");
			value.v = (int)(a + 0.5);
			return true;
		}
		return false;
	}
	static handle cast(PInt s, return_value_policy, handle) { return PyLong_FromLong(s.v); }
};
}}  // namespace pybind11::detail

struct V3 { float x = 0, y = 0, z = 0; };
struct V4 { float x = 0, y = 0, z = 0, t = 0; };

static V4 toV4(const py::handle &o)
{
	V4 r;
	if (py::isinstance<V4>(o)) return o.cast<V4>();
	if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) {
		const float v = o.cast<float>();
		r.x = r.y = r.z = r.t = v;
		return r;
	}
	if (py::isinstance<py::sequence>(o)) {
		py::sequence s = py::reinterpret_borrow<py::sequence>(o);
		if (s.size() != 4) errMsg("argument is not a Vec4");
		r.x = s[0].cast<float>(); r.y = s[1].cast<float>(); r.z = s[2].cast<float>(); r.t = s[3].cast<float>();
		return r;
	}
	errMsg("argument is not a Vec4");
	return r;
}
static V3 toV3(const py::handle &o)
{
	V3 r;
	if (py::isinstance<V3>(o)) return o.cast<V3>();
	if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) {
		r.x = r.y = r.z = o.cast<float>();
		return r;
	}
	if (py::isinstance<py::sequence>(o)) {
		py::sequence s = py::reinterpret_borrow<py::sequence>(o);
		if (s.size() != 3) errMsg("argument is not a Vec3");
		r.x = s[0].cast<float>(); r.y = s[1].cast<float>(); r.z = s[2].cast<float>();
		return r;
	}
	errMsg("argument is not a Vec3");
	return r;
}
static void v4arr(const V4 &v, float o[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.t; }

// ------------------------------------------------------------------ solver + grids --------
// ref: FluidSolver fluidsolver.h:27-119.  The grid pool is the device's stream-ordered pool.
struct Solver {
	std::string name;
	int gx, gy, gz, dim, fourthDim;
	float dt = 1.0f, timeTotal = 0.f;
	int frame = 0;
	Solver(const py::object &gridSize, int dim_, int fourthDim_, const std::string &name_)
	    : name(name_), dim(dim_), fourthDim(fourthDim_)
	{
		const V3 s = toV3(gridSize);
		gx = (int)s.x; gy = (int)s.y; gz = (int)s.z;
		if (!(dim == 2 || dim == 3)) errMsg("Can only create 2D and 3D solvers");
		ctx();  // fail early without a device
	}
	bool has4D() const { return fourthDim > 0; }
};

enum Kind { K_REAL = 0, K_INT = 1, K_VEC3 = 2, K_VEC4 = 3 };
static int kind_elem(int k) { return k == K_VEC4 ? 4 : (k == K_VEC3 ? 3 : 1); }

struct GridAny {
	Solver *parent;
	int kind, elem;
	void *ptr = nullptr;
	int64_t cells = 0;
	std::string name;
	GridAny(Solver *p, int k) : parent(p), kind(k), elem(kind_elem(k))
	{
		if (!p) errMsg("New class: no parent given -- specify using parent=xxx !");
	}
	virtual ~GridAny()
	{
		if (ptr && g_ctx) flof_free(g_ctx, ptr);
	}
	size_t bytes() const { return (size_t)cells * elem * 4; }
	void alloc() { CK(flof_malloc(ctx(), &ptr, bytes()), "Grid"); }
	float *f() const { return (float *)ptr; }
	void sameRes(const GridAny &o, const char *fn) const  // same cell count / dims, element type may differ
	{
		if (o.cells != cells) errMsg(std::string(fn) + ": different grid resolutions");
	}
	void sameSize(const GridAny &o, const char *fn) const
	{
		if (o.cells != cells || o.elem != elem) errMsg(std::string(fn) + ": different grid resolutions / types");
	}
	// element-wise ops shared by 3D and 4D grids (ref grid4d.h:338-372 / grid.h:220-252)
	void clear() { CK(flof_memset0(ctx(), ptr, bytes()), "clear"); }
	void copyFrom(const GridAny &a) { sameSize(a, "copyFrom"); CK(flof_memcpy_d2d(ctx(), ptr, a.ptr, bytes()), "copyFrom"); }
	void add(const GridAny &a) { needFloat("add"); sameSize(a, "add"); CK(flof_grid_binary(ctx(), f(), a.f(), cells, elem, FLOF_OP_ADD), "add"); }
	void sub(const GridAny &a) { needFloat("sub"); sameSize(a, "sub"); CK(flof_grid_binary(ctx(), f(), a.f(), cells, elem, FLOF_OP_SUB), "sub"); }
	void mult(const GridAny &a) { needFloat("mult"); sameSize(a, "mult"); CK(flof_grid_binary(ctx(), f(), a.f(), cells, elem, FLOF_OP_MULT), "mult"); }
	void factor(const py::handle &s, float o[4]) const
	{
		if (kind == K_VEC4) v4arr(toV4(s), o);
		else if (kind == K_VEC3) { const V3 v = toV3(s); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = 0.f; }
		else { o[0] = o[1] = o[2] = o[3] = s.cast<float>(); }
	}
	void needFloat(const char *fn) const { if (kind == K_INT) errMsg(std::string(fn) + ": not supported for int grids on the B200 path"); }
	void elemOp(const char *fn, const py::handle &s, int which, const GridAny *other = nullptr)
	{
		needFloat(fn);
		float o[4];
		factor(s, o);
		if (kind == K_VEC3) {  // 3 floats per cell: treat as a flat float array (component-wise equal factors only)
			if (!(o[0] == o[1] && o[1] == o[2])) errMsg(std::string(fn) + ": per-component factors on Vec3 grids are not on the FlOF path");
			const float q[4] = { o[0], o[0], o[0], o[0] };
			if (which == 0) CK(flof_grid_add_scaled(ctx(), f(), other->f(), cells * 3, 1, q), fn);
			if (which == 1) CK(flof_grid_mult_const(ctx(), f(), cells * 3, 1, q), fn);
			if (which == 2) CK(flof_grid_add_const(ctx(), f(), cells * 3, 1, q), fn);
			if (which == 3) CK(flof_grid_set_const(ctx(), f(), cells * 3, 1, q), fn);
			return;
		}
		if (which == 0) CK(flof_grid_add_scaled(ctx(), f(), other->f(), cells, elem, o), fn);
		if (which == 1) CK(flof_grid_mult_const(ctx(), f(), cells, elem, o), fn);
		if (which == 2) CK(flof_grid_add_const(ctx(), f(), cells, elem, o), fn);
		if (which == 3) CK(flof_grid_set_const(ctx(), f(), cells, elem, o), fn);
	}
	void setConst(const py::object &s)
	{
		if (kind == K_INT) { CK(flof_grid_set_const_int(ctx(), (int *)ptr, cells, s.cast<int>()), "setConst"); return; }
		elemOp("setConst", s, 3);
	}
	void addConst(const py::object &s) { elemOp("addConst", s, 2); }
	void multConst(const py::object &s) { elemOp("multConst", s, 1); }
	void addScaled(const GridAny &a, const py::object &s) { sameSize(a, "addScaled"); elemOp("addScaled", s, 0, &a); }
	void clampv(float lo, float hi)
	{
		needFloat("clamp");
		CK(flof_grid_clamp(ctx(), f(), kind == K_VEC3 ? cells * 3 : cells, kind == K_VEC3 ? 1 : elem, lo, hi), "clamp");
	}
	void minmax(float out[3]) const
	{
		if (kind == K_INT) {
			int mm[2];
			CK(flof_grid_min_max_int(ctx(), (const int *)ptr, cells, mm), "getMax");
			out[0] = (float)mm[0]; out[1] = (float)mm[1];
			out[2] = std::max(std::fabs((float)mm[0]), std::fabs((float)mm[1]));
			return;
		}
		if (kind == K_VEC3) errMsg("getMax: Vec3 grids are not on the FlOF path");
		CK(flof_grid_min_max(ctx(), f(), cells, elem, out), "getMax");
	}
	float getMax() const { float o[3]; minmax(o); return o[1]; }
	float getMin() const { float o[3]; minmax(o); return o[0]; }
	float getMaxAbs() const { float o[3]; minmax(o); return o[2]; }
	std::vector<char> download() const
	{
		std::vector<char> h(bytes());
		CK(flof_memcpy_d2h(ctx(), h.data(), ptr, bytes()), "download");
		return h;
	}
	void upload(const void *h) { CK(flof_memcpy_h2d(ctx(), ptr, h, bytes()), "upload"); CK(flof_sync(ctx()), "upload"); }
};

struct Grid4 : GridAny {  // ref Grid4dBase/Grid4d<T> grid4d.h:26-271
	flof_dim4 d;
	Grid4(Solver *p, int k) : GridAny(p, k)
	{
		if (!(p->dim == 3 && p->has4D())) errMsg("Solver does not support 4d grids , or is two-dimensional!");
		d.nx = p->gx; d.ny = p->gy; d.nz = p->gz; d.nt = p->fourthDim;
		cells = (int64_t)d.nx * d.ny * d.nz * d.nt;
		alloc();
	}
};
struct Grid3 : GridAny {  // ref GridBase/Grid<T> grid.h
	flof_dim3 d;
	Grid3(Solver *p, int k) : GridAny(p, k)
	{
		d.nx = p->gx; d.ny = p->gy; d.nz = p->gz;
		cells = (int64_t)d.nx * d.ny * d.nz;
		alloc();
	}
};
template <int K> struct Grid4T : Grid4 { Grid4T(Solver *p, bool) : Grid4(p, K) {} };
template <int K> struct Grid3T : Grid3 { Grid3T(Solver *p, bool) : Grid3(p, K) {} };
struct LevelsetGrid : Grid3T<K_REAL> { LevelsetGrid(Solver *p, bool s) : Grid3T<K_REAL>(p, s) {} };
struct FlagGrid : Grid3T<K_INT> { FlagGrid(Solver *p, int, bool s) : Grid3T<K_INT>(p, s) {} };
struct Mesh { Solver *parent; Mesh(Solver *p) : parent(p) {} };
struct Timings { };

// ------------------------------------------------------------------ .uni I/O (ref fileio.cpp) ---
#pragma pack(push, 1)
struct UniHeader {  // fileio.cpp:36-43 (288 bytes, natural alignment == packed)
	int dimX, dimY, dimZ;
	int gridType, elementType, bytesPerElement;
	char info[256];
	unsigned long long timestamp;
};
#pragma pack(pop)
static_assert(sizeof(UniHeader) == 288, "UniHeader layout");

// ---- overlapped .uni I/O ----------------------------------------------------------------------
// Mode 3 is I/O bound once the look-up kernel is fast: 180 gz-compressed input slices, 119 gz-compressed output
// frames.  The reference (fileio.cpp:608-700, 834-1002) compresses and decompresses on the calling thread.  Here a
// small worker pool does the zlib work: `save` returns after the device->host copy and the frame is deflated and
// written in the background while the next frame is computed; `loadPlaceGrid4d` inflates its slice files in
// parallel.  Any access to a file with a pending write waits for it; everything is drained at interpreter exit.
class IoPool {
public:
	static IoPool &get()
	{
		static IoPool p;
		return p;
	}
	// runs fn on a worker; `key` (file name, may be empty) lets readers wait for a pending write of that file
	void submit(const std::string &key, size_t bytes, std::function<void()> fn)
	{
		std::unique_lock<std::mutex> lk(mu_);
		if (workers_.empty() || syncIo_) {  // after shutdown (interpreter exit) or FLOF_SYNC_IO=1 (A/B timing): run inline
			lk.unlock();
			fn();
			return;
		}
		cv_done_.wait(lk, [&] { return pendingBytes_ + bytes <= kMaxPendingBytes || pendingBytes_ == 0; });
		if (!key.empty()) pendingKeys_[key]++;
		pendingBytes_ += bytes;
		active_++;
		jobs_.push_back(Job{ key, bytes, std::move(fn) });
		cv_job_.notify_one();
	}
	void waitKey(const std::string &key)
	{
		std::unique_lock<std::mutex> lk(mu_);
		cv_done_.wait(lk, [&] { return pendingKeys_.find(key) == pendingKeys_.end(); });
		rethrow(lk);
	}
	void drain()
	{
		std::unique_lock<std::mutex> lk(mu_);
		cv_done_.wait(lk, [&] { return active_ == 0; });
		rethrow(lk);
	}
	void shutdown()
	{
		{
			std::unique_lock<std::mutex> lk(mu_);
			cv_done_.wait(lk, [&] { return active_ == 0; });
			stop_ = true;
			cv_job_.notify_all();
		}
		for (auto &t : workers_) t.join();
		workers_.clear();
	}
	int threads() const { return (int)workers_.size(); }

private:
	struct Job {
		std::string key;
		size_t bytes;
		std::function<void()> fn;
	};
	static constexpr size_t kMaxPendingBytes = (size_t)2 << 30;
	IoPool()
	{
		syncIo_ = getenv("FLOF_SYNC_IO") != nullptr;
		unsigned n = std::thread::hardware_concurrency();
		n = n == 0 ? 4 : (n > 8 ? 8 : n);
		for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
	}
	~IoPool() { shutdown(); }
	void run()
	{
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_job_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
				if (jobs_.empty()) return;
				j = std::move(jobs_.front());
				jobs_.pop_front();
			}
			std::string err;
			try {
				j.fn();
			} catch (const std::exception &e) {
				err = e.what();
			}
			std::unique_lock<std::mutex> lk(mu_);
			if (!err.empty() && error_.empty()) error_ = err;
			if (!j.key.empty() && --pendingKeys_[j.key] == 0) pendingKeys_.erase(j.key);
			pendingBytes_ -= j.bytes;
			active_--;
			cv_done_.notify_all();
		}
	}
	void rethrow(std::unique_lock<std::mutex> &)
	{
		if (error_.empty()) return;
		const std::string e = error_;
		error_.clear();
		throw std::runtime_error(e);
	}
	std::mutex mu_;
	std::condition_variable cv_job_, cv_done_;
	std::deque<Job> jobs_;
	std::map<std::string, int> pendingKeys_;
	std::vector<std::thread> workers_;
	std::string error_;
	size_t pendingBytes_ = 0;
	int active_ = 0;
	bool stop_ = false;
	bool syncIo_ = false;
};

static void uniSize(const std::string &name, int &x, int &y, int &z, int *t)
{  // ref getUniFileSize fileio.cpp:572-597
	IoPool::get().waitKey(name);
	x = y = z = 0;
	gzFile gzf = gzopen(name.c_str(), "rb");
	if (!gzf) return;
	char ID[5] = { 0, 0, 0, 0, 0 };
	gzread(gzf, ID, 4);
	if (!strcmp(ID, "MNT2") || !strcmp(ID, "M4T2")) {
		UniHeader head;
		if (gzread(gzf, &head, sizeof(head)) != (int)sizeof(head)) { gzclose(gzf); errMsg("can't read file, no header present"); }
		x = head.dimX; y = head.dimY; z = head.dimZ;
	}
	if (!strcmp(ID, "M4T2") && t) {
		int dimT = 0;
		gzread(gzf, &dimT, sizeof(int));
		*t = dimT;
	}
	gzclose(gzf);
}
static void gzReadAll(gzFile gzf, void *dst, size_t n, const std::string &name)
{
	char *p = (char *)dst;
	while (n > 0) {
		const unsigned chunk = (unsigned)std::min<size_t>(n, 1u << 30);
		const int got = gzread(gzf, p, chunk);
		if (got <= 0) { gzclose(gzf); errMsg("can't read file " + name + ": truncated payload"); }
		p += got;
		n -= (size_t)got;
	}
}
static int gridTypeOf(int kind, bool fourd)
{  // GridBase::GridType (grid.h) / Grid4dBase::Grid4dType (grid4d.h:29-35)
	if (fourd) return kind == K_REAL ? 1 : (kind == K_INT ? 2 : (kind == K_VEC3 ? 4 : 8));
	return kind == K_REAL ? 1 : (kind == K_INT ? 2 : 4);
}
static void writeUni(const std::string &name, const GridAny &g, int nx, int ny, int nz, int nt /* 0 = 3D */)
{  // ref writeGridUni :608 / writeGrid4dUni :834
	if (name.find_last_of('.') == std::string::npos) errMsg("file '" + name + "' does not have an extension");
	debMsg(1, (nt ? "writing grid4d " : "Writing grid ") << g.name << " to uni file " << name);
	UniHeader head;
	memset(&head, 0, sizeof(head));
	head.dimX = nx; head.dimY = ny; head.dimZ = nz;
	head.gridType = gridTypeOf(g.kind, nt > 0);
	head.bytesPerElement = g.elem * 4;
	head.elementType = g.kind == K_INT ? 0 : (g.kind == K_REAL ? 1 : 2);
	snprintf(head.info, 256, "%s", "mantaflow flof-b200 (CUDA sm_100a) fp1");
	head.timestamp = 0;
	// device -> host now (the grid may change right after `save` returns); deflate + write on a worker
	auto payload = std::make_shared<std::vector<char>>(g.download());
	const bool fourd = nt > 0;
	IoPool::get().waitKey(name);  // an older pending write of the same file finishes first
	{  // create the file now: unwritable paths fail here like in the reference, and the file exists once save() returns
		FILE *f = fopen(name.c_str(), "wb");
		if (!f) errMsg("can't open file " + name);
		fclose(f);
	}
	IoPool::get().submit(name, payload->size(), [name, head, nt, fourd, payload]() {
		gzFile gzf = gzopen(name.c_str(), "wb1");  // level 1 like the reference (fileio.cpp:862)
		if (!gzf) throw std::runtime_error("can't open file " + name);
		gzwrite(gzf, fourd ? "M4T2" : "MNT2", 4);
		gzwrite(gzf, &head, sizeof(head));
		if (fourd) gzwrite(gzf, &nt, sizeof(int));
		size_t off = 0;
		while (off < payload->size()) {
			const unsigned chunk = (unsigned)std::min<size_t>(payload->size() - off, 1u << 28);
			if (gzwrite(gzf, payload->data() + off, chunk) <= 0) {
				gzclose(gzf);
				throw std::runtime_error("write error on " + name);
			}
			off += chunk;
		}
		if (gzclose(gzf) != Z_OK) throw std::runtime_error("write error on " + name);
	});
}
static void readUni(const std::string &name, GridAny &g, int nx, int ny, int nz, int nt /* 0 = 3D */)
{  // ref readGridUni :648 (MNT2 only) / readGrid4dUni :885 (full grid)
	if (name.find_last_of('.') == std::string::npos) errMsg("file '" + name + "' does not have an extension");
	debMsg(1, "reading grid " << g.name << " from uni file " << name);
	IoPool::get().waitKey(name);
	gzFile gzf = gzopen(name.c_str(), "rb");
	if (!gzf) errMsg("can't open file " + name);
	char ID[5] = { 0, 0, 0, 0, 0 };
	gzread(gzf, ID, 4);
	if (strcmp(ID, nt ? "M4T2" : "MNT2")) { gzclose(gzf); errMsg("Unknown header / legacy uni format in " + name); }
	UniHeader head;
	if (gzread(gzf, &head, sizeof(head)) != (int)sizeof(head)) { gzclose(gzf); errMsg("can't read file, no header present"); }
	if (head.bytesPerElement != g.elem * 4) { gzclose(gzf); errMsg("grid element size doesn't match"); }
	if (!(head.dimX == nx && head.dimY == ny && head.dimZ == nz)) {
		gzclose(gzf);
		std::ostringstream s;
		s << "grid dim doesn't match, [" << head.dimX << "," << head.dimY << "," << head.dimZ << "] vs [" << nx << "," << ny << "," << nz << "]";
		errMsg(s.str());
	}
	if (nt) {
		int fourthDim = 0;
		gzread(gzf, &fourthDim, sizeof(int));
		if (fourthDim != nt) { gzclose(gzf); errMsg("grid dim4 doesn't match"); }
	}
	std::vector<char> h(g.bytes());
	gzReadAll(gzf, h.data(), h.size(), name);
	gzclose(gzf);
	g.upload(h.data());
}

// ------------------------------------------------------------------ plugin helpers --------
static Grid4 &g4(GridAny &g, const char *fn)
{
	Grid4 *p = dynamic_cast<Grid4 *>(&g);
	if (!p) errMsg(std::string(fn) + ": 4D grid expected");
	return *p;
}
template <class T> static T *optGrid(const py::object &o, const char *fn)
{  // optional grid pointer: None or the integer 0 mean NULL (pclass.cpp:116-119)
	if (o.is_none()) return nullptr;
	if (py::isinstance<py::int_>(o) && o.cast<long>() == 0) return nullptr;
	try {
		return o.cast<T *>();
	} catch (py::cast_error &) {
		errMsg(std::string(fn) + ": wrong grid type for optional argument");
	}
	return nullptr;
}
static void requireKind(const GridAny &g, int k, const char *fn, const char *what)
{
	if (g.kind != k) errMsg(std::string(fn) + ": argument '" + what + "' has the wrong grid type");
}

// LATS: per-ID state of the optimised slice look-up (ref LoadAdvectData :1786-1867).  The whole
// deformation volume stays resident on the device instead of a cached gz handle + two slices.
struct Lats {
	static const int MAXDV = 10;  // ref :1781
	flof_dim4 dd;
	void *defo = nullptr;  // deformation volume 0 (whole file)
	std::string fname;
	// defo volumes (`thirdload`, ref :1800-1863): further whole files, one window of Tw slices per volume, the one-slice
	// scratch grid (lats.tmp) and the window-sized scratch of the aligned composition
	bool useDefoVols = false, doAligned = false;
	int Tw = 0, lastT = -1, numDv = 0;
	void *vol[MAXDV] = {}, *win[MAXDV] = {};
	int filepos[MAXDV];
	void *vt = nullptr, *dvt = nullptr;
	void release()
	{
		for (int i = 0; i < MAXDV; ++i) {
			if (vol[i] && vol[i] != defo) flof_free(ctx(), vol[i]);
			if (win[i]) flof_free(ctx(), win[i]);
			vol[i] = win[i] = nullptr;
		}
		if (defo) flof_free(ctx(), defo);
		if (vt) flof_free(ctx(), vt);
		if (dvt) flof_free(ctx(), dvt);
		defo = vt = dvt = nullptr;
		numDv = 0;
	}
};
static std::map<int, Lats> g_lats;
static void *loadDefoVolume(const std::string &fname, const flof_dim4 &dd, const char *fn);

// cache of the hi-res 3D slice sequence of loadPlaceGrid4d, resident on the device
struct SliceSeq {
	flof_dim3 sd;
	int first = 0, count = 0;
	void *data = nullptr;
};
static std::map<std::string, SliceSeq> g_slices;

static SliceSeq &sliceSeq(const std::string &pattern, int first, int end)
{
	SliceSeq &s = g_slices[pattern];
	if (s.data && s.first == first && s.count >= end - first) return s;
	if (s.data) { flof_free(ctx(), s.data); s.data = nullptr; }
	char fn[1024];
	snprintf(fn, sizeof(fn), pattern.c_str(), first);
	int x, y, z;
	uniSize(fn, x, y, z, nullptr);
	if (x < 1 || y < 1 || z < 1) errMsg(std::string("Invalid src size from ") + pattern);
	s.sd.nx = x; s.sd.ny = y; s.sd.nz = z;
	s.first = first;
	s.count = end - first;
	const size_t n3 = (size_t)x * y * z;
	CK(flof_malloc(ctx(), &s.data, n3 * 4 * (size_t)s.count), "loadPlaceGrid4d");
	// inflate the slice files on the I/O workers, a window of them in flight; upload in file order
	const int window = 64;
	for (int b0 = 0; b0 < s.count; b0 += window) {
		const int nb = std::min(window, s.count - b0);
		std::vector<std::vector<float>> bufs((size_t)nb);
		for (int q = 0; q < nb; ++q) {
			snprintf(fn, sizeof(fn), pattern.c_str(), first + b0 + q);
			const std::string name(fn);
			IoPool::get().waitKey(name);
			std::vector<float> *dst = &bufs[(size_t)q];
			IoPool::get().submit(std::string(), n3 * 4, [name, dst, n3, x, y, z]() {
				gzFile gzf = gzopen(name.c_str(), "rb");
				if (!gzf) throw std::runtime_error("can't open file " + name);
				char ID[5] = { 0, 0, 0, 0, 0 };
				gzread(gzf, ID, 4);
				UniHeader head;
				if (strcmp(ID, "MNT2") || gzread(gzf, &head, sizeof(head)) != (int)sizeof(head)) { gzclose(gzf); throw std::runtime_error("bad uni file " + name); }
				if (head.dimX != x || head.dimY != y || head.dimZ != z || head.bytesPerElement != 4) { gzclose(gzf); throw std::runtime_error("grid dim doesn't match in " + name); }
				dst->resize(n3);
				char *p = (char *)dst->data();
				size_t left = n3 * 4;
				while (left > 0) {
					const int got = gzread(gzf, p, (unsigned)std::min<size_t>(left, 1u << 30));
					if (got <= 0) { gzclose(gzf); throw std::runtime_error("can't read file " + name + ": truncated payload"); }
					p += got;
					left -= (size_t)got;
				}
				gzclose(gzf);
			});
		}
		try {
			IoPool::get().drain();
		} catch (const std::exception &e) {
			errMsg(e.what());
		}
		for (int q = 0; q < nb; ++q) CK(flof_memcpy_h2d(ctx(), (float *)s.data + n3 * (size_t)(b0 + q), bufs[(size_t)q].data(), n3 * 4), "loadPlaceGrid4d");
		CK(flof_sync(ctx()), "loadPlaceGrid4d");
	}
	return s;
}

// ------------------------------------------------------------------ module ----------------
template <class G, class... Extra> static void bindCommon(py::class_<G, Extra...> &c)
{
	c.def("clear", [](G &g) { g.clear(); })
	    .def("copyFrom", [](G &g, const GridAny &a, bool) -> G & { g.copyFrom(a); return g; }, py::arg("a"), py::arg("copyType") = true,
	         py::return_value_policy::reference)
	    .def("add", [](G &g, const GridAny &a) { g.add(a); }, py::arg("a"))
	    .def("sub", [](G &g, const GridAny &a) { g.sub(a); }, py::arg("a"))
	    .def("mult", [](G &g, const GridAny &a) { g.mult(a); }, py::arg("a"))
	    .def("setConst", [](G &g, const py::object &s) { g.setConst(s); }, py::arg("s"))
	    .def("addConst", [](G &g, const py::object &s) { g.addConst(s); }, py::arg("s"))
	    .def("multConst", [](G &g, const py::object &s) { g.multConst(s); }, py::arg("s"))
	    .def("addScaled", [](G &g, const GridAny &a, const py::object &f) { g.addScaled(a, f); }, py::arg("a"), py::arg("factor"))
	    .def("clamp", [](G &g, float mn, float mx) { g.clampv(mn, mx); }, py::arg("min"), py::arg("max"))
	    .def("getMax", [](G &g) { return g.getMax(); })
	    .def("getMin", [](G &g) { return g.getMin(); })
	    .def("getMaxAbs", [](G &g) { return g.getMaxAbs(); })
	    .def("getMaxValue", [](G &g) { return g.getMax(); })
	    .def("getMinValue", [](G &g) { return g.getMin(); })
	    .def("getMaxAbsValue", [](G &g) { return g.getMaxAbs(); })
	    .def("setName", [](G &g, const std::string &n) { g.name = n; }, py::arg("name"))
	    .def("getName", [](G &g) { return g.name; });
}

template <class G, class... Extra> static void bind4(py::class_<G, Extra...> &c)
{
	c.def(py::init<Solver *, bool>(), py::arg("parent"), py::arg("show") = true, py::keep_alive<1, 2>());
	bindCommon(c);
	c.def("load", [](G &g, const std::string &name, const py::kwargs &) { readUni(name, g, g.d.nx, g.d.ny, g.d.nz, g.d.nt); }, py::arg("name"))
	    .def("save", [](G &g, const std::string &name, const py::kwargs &) { writeUni(name, g, g.d.nx, g.d.ny, g.d.nz, g.d.nt); }, py::arg("name"))
	    .def("setBound",
	         [](G &g, const py::object &value, PInt boundaryWidth) {
		         if (g.kind == K_INT) { CK(flof_grid4d_set_bound_int(ctx(), (int *)g.ptr, g.d, value.cast<int>(), boundaryWidth.v), "setBound"); return; }
		         if (g.kind == K_VEC3) errMsg("setBound: Grid4Vec3 is not on the FlOF path");
		         float v[4];
		         g.factor(value, v);
		         CK(flof_grid4d_set_bound(ctx(), g.f(), g.d, g.elem, v, boundaryWidth.v), "setBound");
	         },
	         py::arg("value"), py::arg("boundaryWidth") = PInt{ 1 })
	    .def("setBoundNeumann",
	         [](G &g, PInt boundaryWidth) {
		         if (g.kind == K_INT || g.kind == K_VEC3) errMsg("setBoundNeumann: only Real and Vec4 4D grids are on the FlOF path");
		         CK(flof_grid4d_set_bound_neumann(ctx(), g.f(), g.d, g.elem, boundaryWidth.v), "setBoundNeumann");
	         },
	         py::arg("boundaryWidth") = PInt{ 1 })
	    .def("printGrid", [](G &g, PInt, PInt, bool, PInt) { debMsg(1, "Printing '" << g.name << "' (" << g.d.nx << "," << g.d.ny << "," << g.d.nz << "," << g.d.nt << ") max " << g.getMax()); },
	         py::arg("zSlice") = PInt{ -1 }, py::arg("tSlice") = PInt{ -1 }, py::arg("printIndex") = false, py::arg("bnd") = PInt{ 0 })
	    .def("getSize", [](G &g) { V4 v; v.x = (float)g.d.nx; v.y = (float)g.d.ny; v.z = (float)g.d.nz; v.t = (float)g.d.nt; return v; })
	    .def("toNumpyBytes", [](G &g) { const std::vector<char> h = g.download(); return py::bytes(h.data(), h.size()); })
	    .def("fromBytes", [](G &g, const py::bytes &b) { std::string s = b; if (s.size() != g.bytes()) errMsg("fromBytes: size mismatch"); g.upload(s.data()); }, py::arg("data"));
}

template <class G, class... Extra> static void bind3(py::class_<G, Extra...> &c)
{
	bindCommon(c);
	c.def("load", [](G &g, const std::string &name, const py::kwargs &) { readUni(name, g, g.d.nx, g.d.ny, g.d.nz, 0); }, py::arg("name"))
	    .def("save", [](G &g, const std::string &name, const py::kwargs &) { writeUni(name, g, g.d.nx, g.d.ny, g.d.nz, 0); }, py::arg("name"))
	    .def("setBound",
	         [](G &g, const py::object &value, PInt boundaryWidth) {
		         if (g.kind != K_REAL) errMsg("setBound: only Real 3D grids are on the FlOF path");
		         CK(flof_grid3_set_bound(ctx(), g.f(), g.d, value.cast<float>(), boundaryWidth.v), "setBound");
	         },
	         py::arg("value"), py::arg("boundaryWidth") = PInt{ 1 })
	    .def("getSize", [](G &g) { V3 v; v.x = (float)g.d.nx; v.y = (float)g.d.ny; v.z = (float)g.d.nz; return v; })
	    .def("toNumpyBytes", [](G &g) { const std::vector<char> h = g.download(); return py::bytes(h.data(), h.size()); })
	    .def("fromBytes", [](G &g, const py::bytes &b) { std::string s = b; if (s.size() != g.bytes()) errMsg("fromBytes: size mismatch"); g.upload(s.data()); }, py::arg("data"));
}

// a whole Vec4 deformation file resident on the device (ref: the reference streams slices of it through a gz handle)
static void *loadDefoVolume(const std::string &fname, const flof_dim4 &dd, const char *fn)
{
	const size_t bytes = (size_t)dd.nx * dd.ny * dd.nz * dd.nt * 16;
	void *dev = nullptr;
	CK(flof_malloc(ctx(), &dev, bytes), fn);
	IoPool::get().waitKey(fname);
	gzFile gzf = gzopen(fname.c_str(), "rb");
	if (!gzf) { flof_free(ctx(), dev); errMsg("can't open file " + fname); }
	char ID4[5] = { 0, 0, 0, 0, 0 };
	gzread(gzf, ID4, 4);
	UniHeader head;
	int dimT = 0;
	if (strcmp(ID4, "M4T2") || gzread(gzf, &head, sizeof(head)) != (int)sizeof(head) || gzread(gzf, &dimT, 4) != 4) {
		gzclose(gzf);
		flof_free(ctx(), dev);
		errMsg("bad 4d uni file " + fname);
	}
	if (head.bytesPerElement != 16) {
		gzclose(gzf);
		flof_free(ctx(), dev);
		errMsg("grid element size doesn't match (Vec4 deformation expected)");
	}
	std::vector<char> h(bytes);
	gzReadAll(gzf, h.data(), bytes, fname);
	gzclose(gzf);
	CK(flof_memcpy_h2d(ctx(), dev, h.data(), bytes), fn);
	CK(flof_sync(ctx()), fn);
	return dev;
}

PYBIND11_MODULE(manta, m)
{
	m.doc() = "mantaflow-compatible FlOF module on B200 (flof-b200): device-resident Grid4d + CUDA plugin functions";
	m.attr("GUI") = false;
	m.attr("CUDA") = true;
	m.attr("DEBUG") = false;
	m.attr("MT") = false;
	m.attr("DOUBLEPRECISION") = false;
	m.attr("Real") = py::module_::import("builtins").attr("float");
	m.attr("true") = true;
	m.attr("false") = false;
	// flag constants of python/defines.py:24-30
	m.attr("FlagFluid") = 1; m.attr("FlagObstacle") = 2; m.attr("FlagEmpty") = 4; m.attr("FlagStick") = 128;
	m.attr("FlagInflow") = 8; m.attr("FlagOutflow") = 16;

	py::class_<V3>(m, "vec3")
	    .def(py::init([](const py::args &a) {
		    V3 v;
		    if (a.size() == 1) v = toV3(a[0]);
		    else if (a.size() == 3) { v.x = a[0].cast<float>(); v.y = a[1].cast<float>(); v.z = a[2].cast<float>(); }
		    else if (a.size() != 0) errMsg("vec3: 0, 1 or 3 arguments expected");
		    return v;
	    }))
	    .def_readwrite("x", &V3::x).def_readwrite("y", &V3::y).def_readwrite("z", &V3::z)
	    .def("__add__", [](const V3 &a, const py::object &o) { const V3 b = toV3(o); V3 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; return r; })
	    .def("__sub__", [](const V3 &a, const py::object &o) { const V3 b = toV3(o); V3 r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; return r; })
	    .def("__mul__", [](const V3 &a, const py::object &o) { const V3 b = toV3(o); V3 r; r.x = a.x * b.x; r.y = a.y * b.y; r.z = a.z * b.z; return r; })
	    .def("__rmul__", [](const V3 &a, const py::object &o) { const V3 b = toV3(o); V3 r; r.x = a.x * b.x; r.y = a.y * b.y; r.z = a.z * b.z; return r; })
	    .def("__truediv__", [](const V3 &a, const py::object &o) { const V3 b = toV3(o); V3 r; r.x = a.x / b.x; r.y = a.y / b.y; r.z = a.z / b.z; return r; })
	    .def("__len__", [](const V3 &) { return 3; })
	    .def("__getitem__", [](const V3 &a, int i) { if (i < 0 || i > 2) throw py::index_error(); return i == 0 ? a.x : (i == 1 ? a.y : a.z); })
	    .def("__repr__", [](const V3 &a) { char b[128]; snprintf(b, 128, "[%+4.6f,%+4.6f,%+4.6f]", a.x, a.y, a.z); return std::string(b); });
	py::class_<V4>(m, "vec4")
	    .def(py::init([](const py::args &a) {
		    V4 v;
		    if (a.size() == 1) v = toV4(a[0]);
		    else if (a.size() == 4) { v.x = a[0].cast<float>(); v.y = a[1].cast<float>(); v.z = a[2].cast<float>(); v.t = a[3].cast<float>(); }
		    else if (a.size() != 0) errMsg("vec4: 0, 1 or 4 arguments expected");
		    return v;
	    }))
	    .def_readwrite("x", &V4::x).def_readwrite("y", &V4::y).def_readwrite("z", &V4::z).def_readwrite("t", &V4::t)
	    .def("__len__", [](const V4 &) { return 4; })
	    .def("__getitem__", [](const V4 &a, int i) { if (i < 0 || i > 3) throw py::index_error(); return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.t)); })
	    .def("__repr__", [](const V4 &a) { char b[160]; snprintf(b, 160, "[%+4.6f,%+4.6f,%+4.6f,%+4.6f]", a.x, a.y, a.z, a.t); return std::string(b); });
	m.attr("Vec3") = m.attr("vec3");
	m.attr("Vec4") = m.attr("vec4");

	py::class_<Solver>(m, "Solver")
	    .def(py::init([](const py::object &gridSize, PInt dim, PInt fourthDim, const std::string &name) {
		         return new Solver(gridSize, dim.v, fourthDim.v, name);
	         }),
	         py::arg("gridSize"), py::arg("dim") = PInt{ 3 }, py::arg("fourthDim") = PInt{ -1 }, py::arg("name") = "")
	    .def_readwrite("timestep", &Solver::dt)
	    .def_readwrite("timeTotal", &Solver::timeTotal)
	    .def_readwrite("frame", &Solver::frame)
	    .def("getGridSize", [](Solver &s) { V3 v; v.x = (float)s.gx; v.y = (float)s.gy; v.z = (float)s.gz; return v; })
	    .def("printMemInfo", [](Solver &) { debMsg(1, "device-resident grids (stream-ordered CUDA pool)"); })
	    .def("step", [](Solver &s) {  // ref fluidsolver.cpp:168-185 (no GUI): advance time and frame
		         s.timeTotal += s.dt;
		         s.frame++;
		         debMsg(1, s.name << ": frame " << s.frame << " done");
	         })
	    .def("create",
	         [](py::object self, const py::object &type, const py::object &T, const std::string &name) {
		         (void)T;
		         py::object o = type(self);
		         if (!name.empty() && py::hasattr(o, "setName")) o.attr("setName")(name);
		         return o;
	         },
	         py::arg("type"), py::arg("T") = py::none(), py::arg("name") = "");

	py::class_<GridAny>(m, "GridAnyBase");
	py::class_<Grid4, GridAny>(m, "Grid4dBase");
	py::class_<Grid3, GridAny>(m, "GridBase");
	{ py::class_<Grid4T<K_REAL>, Grid4> c(m, "Grid4Real"); bind4(c); }
	{ py::class_<Grid4T<K_VEC4>, Grid4> c(m, "Grid4Vec4"); bind4(c); }
	{ py::class_<Grid4T<K_INT>, Grid4> c(m, "Grid4Int"); bind4(c); }
	{ py::class_<Grid4T<K_VEC3>, Grid4> c(m, "Grid4Vec3"); bind4(c); }
	{
		py::class_<Grid3T<K_REAL>, Grid3> c(m, "RealGrid");
		c.def(py::init<Solver *, bool>(), py::arg("parent"), py::arg("show") = true, py::keep_alive<1, 2>());
		bind3(c);
	}
	{
		py::class_<LevelsetGrid, Grid3T<K_REAL>> c(m, "LevelsetGrid");
		c.def(py::init<Solver *, bool>(), py::arg("parent"), py::arg("show") = true, py::keep_alive<1, 2>());
		c.def("join", [](LevelsetGrid &a, const Grid3 &o) {  // ref levelset.cpp:114-118
			a.sameSize(o, "join");
			CK(flof_grid_binary(ctx(), a.f(), o.f(), a.cells, 1, FLOF_OP_MIN), "join");
		}, py::arg("o"));
		c.def("createMesh", [](LevelsetGrid &, const py::object &, const py::kwargs &) {
			debMsg(2, "createMesh: marching cubes is outside the B200 FlOF path (no-op)");
		}, py::arg("mesh"));
	}
	{
		py::class_<Grid3T<K_VEC3>, Grid3> c(m, "VecGrid");
		c.def(py::init<Solver *, bool>(), py::arg("parent"), py::arg("show") = true, py::keep_alive<1, 2>());
		bind3(c);
		m.attr("Vec3Grid") = m.attr("VecGrid");
		m.attr("MACGrid") = m.attr("VecGrid");  // opticalFlowSimple3d.py passes a MACGrid as the Grid<Vec3> deformation
	}
	{
		py::class_<Grid3T<K_INT>, Grid3> c(m, "IntGrid");
		c.def(py::init<Solver *, bool>(), py::arg("parent"), py::arg("show") = true, py::keep_alive<1, 2>());
		bind3(c);
	}
	{
		py::class_<FlagGrid, Grid3T<K_INT>> c(m, "FlagGrid");
		c.def(py::init<Solver *, int, bool>(), py::arg("parent"), py::arg("dim") = 3, py::arg("show") = true, py::keep_alive<1, 2>());
		// relics in flof.py:339-341 ("setup relics"); the flags are never read on the FlOF path
		c.def("initDomain", [](FlagGrid &, PInt, const py::kwargs &) {}, py::arg("boundaryWidth") = PInt{ 0 });
		c.def("fillGrid", [](FlagGrid &g, PInt type) { CK(flof_grid_set_const_int(ctx(), (int *)g.ptr, g.cells, type.v), "fillGrid"); }, py::arg("type") = PInt{ 1 });
	}
	py::class_<Mesh>(m, "Mesh")
	    .def(py::init<Solver *>(), py::arg("parent"), py::keep_alive<1, 2>())
	    .def("save", [](Mesh &, const std::string &, const py::kwargs &) { errMsg("Mesh.save: triangle meshes are outside the B200 FlOF path"); }, py::arg("name"));
	py::class_<Timings>(m, "Timings").def(py::init<>()).def("display", [](Timings &) {}).def("step", [](Timings &) {});

	// ---------------------------------------------------------------- free functions
	m.def("setDebugLevel", [](PInt level) { g_debug_level = level.v; }, py::arg("level") = PInt{ 1 });
	m.def("printBuildInfo", []() { py::print("mantaflow flof-b200 (CUDA sm_100a) fp1"); return std::string("flof-b200 fp1"); });
	// pending background writes: flushed at interpreter exit, or explicitly
	m.def("flushUniWrites", []() { try { IoPool::get().drain(); } catch (const std::exception &e) { errMsg(e.what()); } });
	py::module_::import("atexit").attr("register")(py::cpp_function([]() {
		try {
			IoPool::get().drain();
		} catch (const std::exception &e) {
			// nobody called flushUniWrites(): a lost .uni write must not look like success to the calling process
			fprintf(stderr, "flof-b200: background .uni write failed: %s\n", e.what());
			fflush(stderr);
			IoPool::get().shutdown();
			_exit(74);  // EX_IOERR
		}
		IoPool::get().shutdown();
	}));
	m.def("getUniFileSize", [](const std::string &name) { int x, y, z; uniSize(name, x, y, z, nullptr); V3 v; v.x = (float)x; v.y = (float)y; v.z = (float)z; return v; }, py::arg("name"));

	// ref opticalFlowMultiscale4d optflow4d.cpp:2182-2195
	m.def("opticalFlowMultiscale4d",
	      [](Grid4 &vel, Grid4 &i0, Grid4 &i1, const py::object &rhsT, float wSmooth, float wEnergy, PInt level, float postVelBlur,
	         float cgAccuracy, PInt blurType, float cfl, PInt orderTime, PInt orderSpace, float resetBndWidth, PInt multiStep,
	         PInt projSizeThresh, PInt minGridSize, bool doFinalProject) {
		      const char *fn = "opticalFlowMultiscale4d";
		      requireKind(vel, K_VEC4, fn, "vel"); requireKind(i0, K_REAL, fn, "i0"); requireKind(i1, K_REAL, fn, "i1");
		      i0.sameSize(i1, fn);
		      if (vel.cells != i0.cells) errMsg(std::string(fn) + ": different grid resolutions");
		      if (vel.parent->dt != 1.0f) errMsg("Invalid, only dt 1 for now!");
		      if (blurType.v != 1) errMsg("NYI");
		      if (level.v != 0) errMsg(std::string(fn) + ": level must be 0 when called from a scene");
		      if (optGrid<Grid4>(rhsT, fn)) errMsg(std::string(fn) + ": rhsT is not filled by the multi-scale driver on the B200 path (use opticalFlow4d)");
		      (void)orderTime; (void)orderSpace;
		      flof_multiscale_params p;
		      flof_multiscale_defaults(&p);
		      p.wSmooth = wSmooth; p.wEnergy = wEnergy; p.postVelBlur = postVelBlur; p.cgAccuracy = cgAccuracy; p.cfl = cfl;
		      p.resetBndWidth = resetBndWidth; p.multiStep = multiStep.v; p.projSizeThresh = projSizeThresh.v;
		      p.minGridSize = minGridSize.v; p.doFinalProject = doFinalProject ? 1 : 0;
		      flof_multiscale_trace tr;
		      float err = 0.f;
		      debMsg(1, "Solving FlOF [" << vel.d.nx << "," << vel.d.ny << "," << vel.d.nz << "," << vel.d.nt << "] on B200");
		      CK(flof_optical_flow_multiscale4d(ctx(), vel.f(), i0.f(), i1.f(), vel.d, &p, &tr, &err), fn);
		      for (int q = 0; q < tr.n_solves && q < 64; ++q)
			      debMsg(1, "ofSolve fix iterations:" << tr.cg_iters[q] << " (" << tr.cg_ms[q] / 1000.f << "s, " << tr.cg_cells[q] << " cells) ");
		      for (int q = 0; q + 1 < tr.n_errs && q < 63; ++q) debMsg(1, "Current error, step " << q << " = " << tr.errs[q]);
		      debMsg(1, "Final error=" << err << "  (device time " << tr.total_ms / 1000.f << "s)");
	      },
	      py::arg("vel"), py::arg("i0"), py::arg("i1"), py::arg("rhsT") = py::none(), py::arg("wSmooth") = 0.f, py::arg("wEnergy") = 0.f,
	      py::arg("level") = PInt{ 0 }, py::arg("postVelBlur") = 0.f, py::arg("cgAccuracy") = 1e-04f, py::arg("blurType") = PInt{ 1 },
	      py::arg("cfl") = 999.f, py::arg("orderTime") = PInt{ 1 }, py::arg("orderSpace") = PInt{ 1 }, py::arg("resetBndWidth") = -1.f,
	      py::arg("multiStep") = PInt{ 1 }, py::arg("projSizeThresh") = PInt{ 9999 }, py::arg("minGridSize") = PInt{ 10 },
	      py::arg("doFinalProject") = false);

	// ---- the 3D instantiations (scenes/opticalFlowSimple3d.py): Grid<Real> / Grid<Vec3> on a 3D solver ----
	// ref opticalFlowMultiscale3d optflow4d.cpp:1175-1188
	m.def("opticalFlowMultiscale3d",
	      [](Grid3 &vel, Grid3 &i0, Grid3 &i1, const py::object &rhsT, float wSmooth, float wEnergy, PInt level, float postVelBlur,
	         float cgAccuracy, PInt blurType, float cfl, PInt orderTime, PInt orderSpace, float resetBndWidth, PInt multiStep,
	         PInt projSizeThresh, PInt minGridSize, bool doFinalProject) {
		      const char *fn = "opticalFlowMultiscale3d";
		      requireKind(vel, K_VEC3, fn, "vel"); requireKind(i0, K_REAL, fn, "i0"); requireKind(i1, K_REAL, fn, "i1");
		      i0.sameSize(i1, fn); vel.sameRes(i0, fn);
		      if (vel.parent->dt != 1.0f) errMsg("Invalid, only dt 1 for now!");
		      if (blurType.v != 1) errMsg("NYI");
		      if (level.v != 0) errMsg(std::string(fn) + ": level must be 0 when called from a scene");
		      if (optGrid<Grid3>(rhsT, fn)) errMsg(std::string(fn) + ": rhsT is not filled by the multi-scale driver on the B200 path");
		      (void)orderTime; (void)orderSpace;
		      flof_multiscale_params p;
		      flof_multiscale_defaults(&p);
		      p.wSmooth = wSmooth; p.wEnergy = wEnergy; p.postVelBlur = postVelBlur; p.cgAccuracy = cgAccuracy; p.cfl = cfl;
		      p.resetBndWidth = resetBndWidth; p.multiStep = multiStep.v; p.projSizeThresh = projSizeThresh.v;
		      p.minGridSize = minGridSize.v; p.doFinalProject = doFinalProject ? 1 : 0;
		      flof_multiscale_trace tr;
		      float err = 0.f;
		      debMsg(1, "Solving FlOF [" << vel.d.nx << "," << vel.d.ny << "," << vel.d.nz << "] on B200");
		      CK(flof_optical_flow_multiscale3d(ctx(), vel.f(), i0.f(), i1.f(), vel.d, &p, &tr, &err), fn);
		      for (int q = 0; q < tr.n_solves && q < 64; ++q) debMsg(1, "ofSolve fix iterations:" << tr.cg_iters[q] << " ");
		      for (int q = 0; q + 1 < tr.n_errs && q < 63; ++q) debMsg(1, "Current error, step " << q << " = " << tr.errs[q]);
		      debMsg(1, "Final error=" << err << "  (device time " << tr.total_ms / 1000.f << "s)");
	      },
	      py::arg("vel"), py::arg("i0"), py::arg("i1"), py::arg("rhsT") = py::none(), py::arg("wSmooth") = 0.f, py::arg("wEnergy") = 0.f,
	      py::arg("level") = PInt{ 0 }, py::arg("postVelBlur") = 0.f, py::arg("cgAccuracy") = 1e-04f, py::arg("blurType") = PInt{ 1 },
	      py::arg("cfl") = 999.f, py::arg("orderTime") = PInt{ 1 }, py::arg("orderSpace") = PInt{ 1 }, py::arg("resetBndWidth") = -1.f,
	      py::arg("multiStep") = PInt{ 1 }, py::arg("projSizeThresh") = PInt{ 9999 }, py::arg("minGridSize") = PInt{ 10 },
	      py::arg("doFinalProject") = false);
	// ref corrVelsOf3d :803-812
	m.def("corrVelsOf3d",
	      [](Grid3 &dst, Grid3 &vel, Grid3 &phiOrg, Grid3 &phiCurr, Grid3 &phiTarget, float threshPhi, float threshNorm, float postVelBlur,
	         float resetBndWidth, PInt maxIter) {
		      (void)threshNorm;
		      const char *fn = "corrVelsOf3d";
		      requireKind(dst, K_VEC3, fn, "dst"); requireKind(vel, K_VEC3, fn, "vel"); requireKind(phiOrg, K_REAL, fn, "phiOrg");
		      requireKind(phiCurr, K_REAL, fn, "phiCurr"); requireKind(phiTarget, K_REAL, fn, "phiTarget");
		      vel.sameRes(dst, fn); vel.sameRes(phiOrg, fn); vel.sameRes(phiTarget, fn);
		      CK(flof_corr_vels_of3d(ctx(), dst.f(), vel.f(), phiOrg.f(), phiTarget.f(), vel.d, threshPhi, postVelBlur, resetBndWidth, maxIter.v), fn);
	      },
	      py::arg("dst"), py::arg("vel"), py::arg("phiOrg"), py::arg("phiCurr"), py::arg("phiTarget"), py::arg("threshPhi") = 1e10f,
	      py::arg("threshNorm") = 1e10f, py::arg("postVelBlur") = 0.f, py::arg("resetBndWidth") = -1.f, py::arg("maxIter") = PInt{ 100 });
	// ref advectSemiLagrangeCfl :863-872 (flags, order and orderSpace are unused by the reference too), advectCent3d :836-846
	auto advect3 = [](const char *fn, float cfl, Grid3 &vel, Grid3 &grid, float velFactor) {
		requireKind(vel, K_VEC3, fn, "vel");
		if (grid.kind != K_REAL && grid.kind != K_VEC3) errMsg("AdvectSemiLagrange3d: Grid Type is not supported (only Real, Vec3)");
		vel.sameRes(grid, fn);
		if (vel.parent->dt != 1.0f) errMsg(std::string(fn) + ": only dt 1 on the B200 path");
		CK(flof_advect_semi_lagrange_cfl3d(ctx(), cfl, vel.f(), grid.f(), grid.elem, vel.d, velFactor), fn);
	};
	m.def("advectSemiLagrangeCfl",
	      [advect3](float cfl, Grid3 &flags, Grid3 &vel, Grid3 &grid, PInt order, float velFactor, PInt orderSpace) {
		      (void)flags; (void)order; (void)orderSpace;
		      advect3("advectSemiLagrangeCfl", cfl, vel, grid, velFactor);
	      },
	      py::arg("cfl"), py::arg("flags"), py::arg("vel"), py::arg("grid"), py::arg("order") = PInt{ 1 }, py::arg("velFactor") = 1.f,
	      py::arg("orderSpace") = PInt{ 1 });
	m.def("advectCent3d", [advect3](Grid3 &vel, Grid3 &grid) { advect3("advectCent3d", 3.0e38f, vel, grid, 1.f); }, py::arg("vel"),
	      py::arg("grid"));
	// ref calcLsDiff3d :928-933
	m.def("calcLsDiff3d",
	      [](Grid3 &i0, Grid3 &i1, const py::object &out, float correction, PInt bnd) {
		      Grid3 *o = optGrid<Grid3>(out, "calcLsDiff3d");
		      requireKind(i0, K_REAL, "calcLsDiff3d", "i0"); requireKind(i1, K_REAL, "calcLsDiff3d", "i1"); i0.sameSize(i1, "calcLsDiff3d");
		      if (o) { requireKind(*o, K_REAL, "calcLsDiff3d", "out"); i0.sameSize(*o, "calcLsDiff3d"); }
		      float r = 0.f;
		      CK(flof_calc_ls_diff3d(ctx(), i0.f(), i1.f(), o ? o->f() : nullptr, i0.d, correction, bnd.v, &r), "calcLsDiff3d");
		      return r;
	      },
	      py::arg("i0"), py::arg("i1"), py::arg("out") = py::none(), py::arg("correction") = 1.f, py::arg("bnd") = PInt{ 0 });

	// ref opticalFlow4d :2110
	m.def("opticalFlow4d",
	      [](Grid4 &vel, Grid4 &i0, Grid4 &i1, const py::object &rhsT, float wSmooth, float wEnergy, float postVelBlur, float cgAccuracy,
	         PInt blurType, float resetBndWidth) {
		      const char *fn = "opticalFlow4d";
		      requireKind(vel, K_VEC4, fn, "vel"); requireKind(i0, K_REAL, fn, "i0"); requireKind(i1, K_REAL, fn, "i1");
		      vel.sameRes(i0, fn); vel.sameRes(i1, fn);
		      if (blurType.v != 1) errMsg("NYI");
		      Grid4 *r = optGrid<Grid4>(rhsT, fn);
		      if (r) { requireKind(*r, K_REAL, fn, "rhsT"); vel.sameRes(*r, fn); }
		      int it = 0; float res = 0.f;
		      CK(flof_optical_flow4d(ctx(), vel.f(), i0.f(), i1.f(), r ? r->f() : nullptr, vel.d, wSmooth, wEnergy, postVelBlur, cgAccuracy,
		                             resetBndWidth, &it, &res), fn);
		      debMsg(1, "ofSolve fix iterations:" << it << " ");
	      },
	      py::arg("vel"), py::arg("i0"), py::arg("i1"), py::arg("rhsT") = py::none(), py::arg("wSmooth") = 0.f, py::arg("wEnergy") = 0.f,
	      py::arg("postVelBlur") = 0.f, py::arg("cgAccuracy") = 1e-04f, py::arg("blurType") = PInt{ 1 }, py::arg("resetBndWidth") = -1.f);

	// ref corrVelsOf4d :2121
	m.def("corrVelsOf4d",
	      [](Grid4 &dst, Grid4 &vel, Grid4 &phiOrg, Grid4 &phiCurr, Grid4 &phiTarget, float threshPhi, float threshNorm, float postVelBlur,
	         float resetBndWidth, PInt maxIter) {
		      (void)threshNorm;
		      const char *fn = "corrVelsOf4d";
		      requireKind(dst, K_VEC4, fn, "dst"); requireKind(vel, K_VEC4, fn, "vel"); requireKind(phiOrg, K_REAL, fn, "phiOrg");
		      requireKind(phiCurr, K_REAL, fn, "phiCurr"); requireKind(phiTarget, K_REAL, fn, "phiTarget");
		      vel.sameRes(dst, fn); vel.sameRes(phiOrg, fn); vel.sameRes(phiTarget, fn);
		      CK(flof_corr_vels_of4d(ctx(), dst.f(), vel.f(), phiOrg.f(), phiTarget.f(), vel.d, threshPhi, postVelBlur, resetBndWidth, maxIter.v),
		         "corrVelsOf4d");
	      },
	      py::arg("dst"), py::arg("vel"), py::arg("phiOrg"), py::arg("phiCurr"), py::arg("phiTarget"), py::arg("threshPhi") = 1e10f,
	      py::arg("threshNorm") = 1e10f, py::arg("postVelBlur") = 0.f, py::arg("resetBndWidth") = -1.f, py::arg("maxIter") = PInt{ 100 });

	// ref calcLsDiff4d :2132 / calcSmokeDiff4d :2163
	m.def("calcLsDiff4d",
	      [](Grid4 &i0, Grid4 &i1, const py::object &out, float correction, PInt bnd) {
		      Grid4 *o = optGrid<Grid4>(out, "calcLsDiff4d");
		      requireKind(i0, K_REAL, "calcLsDiff4d", "i0"); requireKind(i1, K_REAL, "calcLsDiff4d", "i1"); i0.sameSize(i1, "calcLsDiff4d");
		      if (o) { requireKind(*o, K_REAL, "calcLsDiff4d", "out"); i0.sameSize(*o, "calcLsDiff4d"); }
		      float r = 0.f;
		      CK(flof_calc_ls_diff4d(ctx(), i0.f(), i1.f(), o ? o->f() : nullptr, i0.d, correction, bnd.v, &r), "calcLsDiff4d");
		      return r;
	      },
	      py::arg("i0"), py::arg("i1"), py::arg("out") = py::none(), py::arg("correction") = 1.f, py::arg("bnd") = PInt{ 0 });
	m.def("calcSmokeDiff4d",
	      [](Grid4 &i0, Grid4 &i1, const py::object &, float correction, PInt bnd) {
		      requireKind(i0, K_REAL, "calcSmokeDiff4d", "i0"); requireKind(i1, K_REAL, "calcSmokeDiff4d", "i1"); i0.sameSize(i1, "calcSmokeDiff4d");
		      float r = 0.f;
		      CK(flof_calc_smoke_diff4d(ctx(), i0.f(), i1.f(), i0.d, correction, bnd.v, &r), "calcSmokeDiff4d");
		      return r;
	      },
	      py::arg("i0"), py::arg("i1"), py::arg("out") = py::none(), py::arg("correction") = 1.f, py::arg("bnd") = PInt{ 0 });

	// ref advect4d :1292 (dt = solver dt * dtFac)
	m.def("advect4d",
	      [](Grid4 &vel, Grid4 &grid, float dtFac) {
		      requireKind(vel, K_VEC4, "advect4d", "vel");
		      if (!(grid.kind == K_REAL || grid.kind == K_VEC4)) errMsg("AdvectSemiLagrange4d: Grid Type is not supported (only Real, Vec4 on B200)");
		      vel.sameRes(grid, "advect4d");
		      CK(flof_advect4d(ctx(), vel.f(), grid.f(), grid.d, grid.elem, vel.parent->dt * dtFac), "advect4d");
	      },
	      py::arg("vel"), py::arg("grid"), py::arg("dtFac") = 1.f);

	// ref extrap4dLsSimple :1361, extrapolateVec4Simple :1408, repeatFrame4d :1254
	m.def("extrap4dLsSimple", [](Grid4 &phi, PInt distance, bool inside) {
		      requireKind(phi, K_REAL, "extrap4dLsSimple", "phi");
		      CK(flof_extrap4d_ls_simple(ctx(), phi.f(), phi.d, distance.v, inside ? 1 : 0, nullptr), "extrap4dLsSimple");
	      }, py::arg("phi"), py::arg("distance") = PInt{ 4 }, py::arg("inside") = false);
	m.def("extrapolateVec4Simple", [](Grid4 &vel, Grid4 &phi, PInt distance) {
		      requireKind(vel, K_VEC4, "extrapolateVec4Simple", "vel"); requireKind(phi, K_REAL, "extrapolateVec4Simple", "phi");
		      vel.sameRes(phi, "extrapolateVec4Simple");
		      CK(flof_extrapolate_vec4_simple(ctx(), vel.f(), phi.f(), phi.d, distance.v), "extrapolateVec4Simple");
	      }, py::arg("vel"), py::arg("phi"), py::arg("distance"));
	m.def("repeatFrame4d", [](Grid4 &phi, float srct, float range, PInt bnd) {
		      requireKind(phi, K_REAL, "repeatFrame4d", "phi");
		      CK(flof_repeat_frame4d(ctx(), phi.f(), phi.d, srct, range, bnd.v), "repeatFrame4d");
	      }, py::arg("phi"), py::arg("srct"), py::arg("range") = 0.f, py::arg("bnd") = PInt{ 0 });

	// ref interpolateGrid4d / interpolateGrid4dVec grid4d.cpp:539-557
	auto interp = [](Grid4 &target, Grid4 &source, const py::object &offset, const py::object &scale, const py::object &size, const char *fn) {
		if (target.kind != source.kind) errMsg(std::string(fn) + ": grid types differ");
		float o[4], s[4], z[4];
		v4arr(toV4(offset), o); v4arr(toV4(scale), s); v4arr(toV4(size), z);
		CK(flof_interpolate_grid4d(ctx(), target.f(), target.d, source.f(), source.d, source.elem, o, s, z), fn);
	};
	m.def("interpolateGrid4d", [interp](Grid4 &t, Grid4 &s, const py::object &o, const py::object &sc, const py::object &sz) {
		      requireKind(t, K_REAL, "interpolateGrid4d", "target"); interp(t, s, o, sc, sz, "interpolateGrid4d");
	      }, py::arg("target"), py::arg("source"), py::arg("offset") = py::float_(0.), py::arg("scale") = py::float_(1.), py::arg("size") = py::float_(-1.));
	m.def("interpolateGrid4dVec", [interp](Grid4 &t, Grid4 &s, const py::object &o, const py::object &sc, const py::object &sz) {
		      requireKind(t, K_VEC4, "interpolateGrid4dVec", "target"); interp(t, s, o, sc, sz, "interpolateGrid4dVec");
	      }, py::arg("target"), py::arg("source"), py::arg("offset") = py::float_(0.), py::arg("scale") = py::float_(1.), py::arg("size") = py::float_(-1.));

	// ref slices / components / regions grid4d.cpp:338-353, 466-524
	m.def("getSliceFrom4d", [](Grid4 &src, PInt srct, Grid3 &dst) {
		      if (dst.d.nx != src.d.nx || dst.d.ny != src.d.ny || dst.d.nz != src.d.nz) errMsg("getSliceFrom4d: 3D size of dst must match src");
		      CK(flof_get_slice_from4d(ctx(), src.f(), src.d, srct.v, dst.f()), "getSliceFrom4d");
	      }, py::arg("src"), py::arg("srct"), py::arg("dst"));
	m.def("getSliceFrom4dVec", [](Grid4 &src, PInt srct, Grid3 &dst, const py::object &dstt) {
		      Grid3 *tt = optGrid<Grid3>(dstt, "getSliceFrom4dVec");
		      if (dst.d.nx != src.d.nx || dst.d.ny != src.d.ny || dst.d.nz != src.d.nz) errMsg("getSliceFrom4dVec: 3D size of dst must match src");
		      CK(flof_get_slice_from4d_vec(ctx(), src.f(), src.d, srct.v, dst.f(), tt ? tt->f() : nullptr), "getSliceFrom4dVec");
	      }, py::arg("src"), py::arg("srct"), py::arg("dst"), py::arg("dstt") = py::none());
	m.def("placeGrid3d", [](Grid3 &src, Grid4 &dst, PInt dstt) {
		      if (src.d.nx != dst.d.nx || src.d.ny != dst.d.ny || src.d.nz != dst.d.nz) errMsg("placeGrid3d: 3D size of src must match dst");
		      CK(flof_place_grid3d(ctx(), src.f(), dst.f(), dst.d, dstt.v), "placeGrid3d");
	      }, py::arg("src"), py::arg("dst"), py::arg("dstt"));
	m.def("getComp4d", [](Grid4 &src, Grid4 &dst, PInt c) { requireKind(src, K_VEC4, "getComp4d", "src"); requireKind(dst, K_REAL, "getComp4d", "dst"); src.sameRes(dst, "getComp4d"); CK(flof_get_comp4d(ctx(), src.f(), dst.f(), src.cells, c.v), "getComp4d"); }, py::arg("src"), py::arg("dst"), py::arg("c"));
	m.def("setComp4d", [](Grid4 &src, Grid4 &dst, PInt c) { requireKind(src, K_REAL, "setComp4d", "src"); requireKind(dst, K_VEC4, "setComp4d", "dst"); src.sameRes(dst, "setComp4d"); CK(flof_set_comp4d(ctx(), src.f(), dst.f(), src.cells, c.v), "setComp4d"); }, py::arg("src"), py::arg("dst"), py::arg("c"));
	m.def("setRegion4d", [](Grid4 &dst, const py::object &start, const py::object &end, float value) {
		      float s[4], e[4]; v4arr(toV4(start), s); v4arr(toV4(end), e);
		      const float v[4] = { value, value, value, value };
		      CK(flof_set_region4d(ctx(), dst.f(), dst.d, 1, s, e, v), "setRegion4d");
	      }, py::arg("dst"), py::arg("start"), py::arg("end"), py::arg("value"));
	m.def("setRegion4dVec4", [](Grid4 &dst, const py::object &start, const py::object &end, const py::object &value) {
		      float s[4], e[4], v[4]; v4arr(toV4(start), s); v4arr(toV4(end), e); v4arr(toV4(value), v);
		      CK(flof_set_region4d(ctx(), dst.f(), dst.d, 4, s, e, v), "setRegion4dVec4");
	      }, py::arg("dst"), py::arg("start"), py::arg("end"), py::arg("value"));
	auto maxDiff = [](Grid4 &a, Grid4 &b) { double o = 0.; a.sameSize(b, "grid4dMaxDiff"); CK(flof_grid_max_diff(ctx(), a.f(), b.f(), a.cells, a.elem, &o), "grid4dMaxDiff"); return (float)o; };
	m.def("grid4dMaxDiff", maxDiff, py::arg("g1"), py::arg("g2"));
	m.def("grid4dMaxDiffVec4", maxDiff, py::arg("g1"), py::arg("g2"));
	m.def("grid4dMaxDiffVec3", maxDiff, py::arg("g1"), py::arg("g2"));  // ref grid4d.cpp:439-451 (elem 3: sum of |component diffs|)
	m.def("grid4dMaxDiffInt", [](Grid4 &a, Grid4 &b) {  // ref grid4d.cpp:429-437
		      double o = 0.; a.sameSize(b, "grid4dMaxDiffInt"); requireKind(a, K_INT, "grid4dMaxDiffInt", "g1");
		      CK(flof_grid_max_diff(ctx(), a.f(), b.f(), a.cells, -1, &o), "grid4dMaxDiffInt"); return (float)o;
	      }, py::arg("g1"), py::arg("g2"));
	m.def("debugVelAvg4d", [](Grid4 &v, PInt brd) {  // ref test.cpp:210
		      requireKind(v, K_VEC4, "debugVelAvg4d", "v"); float o = 0.f;
		      CK(flof_debug_vel_avg4d(ctx(), v.f(), v.d, brd.v, &o), "debugVelAvg4d"); return o;
	      }, py::arg("v"), py::arg("brd") = PInt{ 0 });
	m.def("calcObfDiff", [](Grid3 &phi1, Grid3 &phi2, Grid3 &phiDiff, Grid3 &vel1, Grid3 &vel2, Grid3 &velt1, Grid3 &velt2, Grid3 &velDiff, PInt bnd) {
		      requireKind(phi1, K_REAL, "calcObfDiff", "phi1"); requireKind(phi2, K_REAL, "calcObfDiff", "phi2");
		      requireKind(phiDiff, K_REAL, "calcObfDiff", "phiDiff"); requireKind(vel1, K_VEC3, "calcObfDiff", "vel1");
		      requireKind(vel2, K_VEC3, "calcObfDiff", "vel2"); requireKind(velt1, K_REAL, "calcObfDiff", "velt1");
		      requireKind(velt2, K_REAL, "calcObfDiff", "velt2"); requireKind(velDiff, K_REAL, "calcObfDiff", "velDiff");
		      phi1.sameSize(phi2, "calcObfDiff"); phi1.sameSize(phiDiff, "calcObfDiff"); vel1.sameSize(vel2, "calcObfDiff");
		      CK(flof_calc_obf_diff(ctx(), phi1.f(), phi2.f(), phiDiff.f(), vel1.f(), vel2.f(), velt1.f(), velt2.f(), velDiff.f(), phiDiff.d, bnd.v), "calcObfDiff");
	      }, py::arg("phi1"), py::arg("phi2"), py::arg("phiDiff"), py::arg("vel1"), py::arg("vel2"), py::arg("velt1"), py::arg("velt2"),
	      py::arg("velDiff"), py::arg("bnd"));
	m.def("debugGridAvg4d", [](Grid4 &phi, PInt brd) { float o = 0.f; CK(flof_debug_grid_avg4d(ctx(), phi.f(), phi.d, brd.v, &o), "debugGridAvg4d"); return o; }, py::arg("phi"), py::arg("brd") = PInt{ 0 });
	m.def("initVecFromScalar", [](Grid4 &source, Grid4 &target) { requireKind(source, K_REAL, "initVecFromScalar", "source"); requireKind(target, K_VEC4, "initVecFromScalar", "target"); source.sameRes(target, "initVecFromScalar"); CK(flof_init_vec_from_scalar(ctx(), source.f(), target.f(), source.cells), "initVecFromScalar"); }, py::arg("source"), py::arg("target"));
	m.def("initTestCheckerboard", [](Grid4 &val, const py::object &vec, PInt brd) {
		      Grid4 *v = optGrid<Grid4>(vec, "initTestCheckerboard");
		      CK(flof_init_test_checkerboard(ctx(), val.f(), v ? v->f() : nullptr, val.d, brd.v), "initTestCheckerboard");
	      }, py::arg("val"), py::arg("vec") = py::none(), py::arg("brd") = PInt{ 0 });

	// ref simpleBlurSpecial test.cpp:127
	m.def("simpleBlurSpecial", [](Grid3 &a, PInt iter, float thresh, PInt bord) {
		      requireKind(a, K_REAL, "simpleBlurSpecial", "a");
		      CK(flof_simple_blur_special(ctx(), a.f(), a.d, iter.v, thresh, bord.v), "simpleBlurSpecial");
	      }, py::arg("a"), py::arg("iter") = PInt{ 1 }, py::arg("thresh") = 0.f, py::arg("bord") = PInt{ 0 });
	m.def("projectPpmFull", [](Grid3 &, const std::string &, PInt, float) { debMsg(2, "projectPpmFull: image output is outside the B200 FlOF path (no-op)"); },
	      py::arg("val"), py::arg("name"), py::arg("shadeMode") = PInt{ 0 }, py::arg("scale") = 1.f);

	// ref loadPlaceGrid4d :1464-1595 (the 3D slice files are read once and kept on the device)
	m.def("loadPlaceGrid4d",
	      [](const std::string &fname, Grid4 &phi, const py::object &offset, const py::object &scale, PInt fileIdxStart, PInt fileIdxEnd,
	         PInt debugSkipLoad, float spread, const py::object &overrideSize, float overrideTimeOff, PInt overrideGoodRegion,
	         float loadTimeScale, bool rescaleSdfValues, float sdfIsoOff, float repeatStartFrame) {
		      const char *fn = "loadPlaceGrid4d";
		      requireKind(phi, K_REAL, fn, "phi");
		      float o[4], s[4], z[4];
		      v4arr(toV4(offset), o); v4arr(toV4(scale), s); v4arr(toV4(overrideSize), z);
		      int fs = fileIdxStart.v, fe = fileIdxEnd.v;
		      const float defoT = z[0] > 0.f ? z[3] : (float)phi.d.nt;
		      if (fs < 0) fs = 0;
		      if (fe < 0) fe = (int)defoT;
		      const int fend = fe < debugSkipLoad.v ? fe : debugSkipLoad.v;
		      if (fend <= fs) errMsg(std::string(fn) + ": empty file range");
		      SliceSeq &seq = sliceSeq(fname, fs, fend);
		      debMsg(1, "Found size [" << seq.sd.nx << "," << seq.sd.ny << "," << seq.sd.nz << "] in " << fname << " (" << seq.count << " slices resident)");
		      CK(flof_load_place_grid4d(ctx(), (const float *)seq.data, seq.count, seq.sd, phi.f(), phi.d, o, s, fs, fe, debugSkipLoad.v, spread, z,
		                                overrideTimeOff, overrideGoodRegion.v, loadTimeScale, rescaleSdfValues ? 1 : 0, sdfIsoOff, repeatStartFrame), fn);
	      },
	      py::arg("fname"), py::arg("phi"), py::arg("offset"), py::arg("scale"), py::arg("fileIdxStart") = PInt{ -1 }, py::arg("fileIdxEnd") = PInt{ -1 },
	      py::arg("debugSkipLoad") = PInt{ 999999 }, py::arg("spread") = 1.f, py::arg("overrideSize") = py::float_(-1.), py::arg("overrideTimeOff") = 0.f,
	      py::arg("overrideGoodRegion") = PInt{ 0 }, py::arg("loadTimeScale") = 1.f, py::arg("rescaleSdfValues") = false, py::arg("sdfIsoOff") = 0.f,
	      py::arg("repeatStartFrame") = 0.f);
	m.def("shiftForwGrid4d", [](Grid4 &phi, PInt overrideGoodRegion) { CK(flof_shift_forw_grid4d(ctx(), phi.f(), phi.d, overrideGoodRegion.v), "shiftForwGrid4d"); },
	      py::arg("phi"), py::arg("overrideGoodRegion") = PInt{ 0 });

	// ref loadAdvectTimeSlice_OptInit :1871, _OptAdd :1914, _Finish :1930, _OptRun :1951, loadAdvectTimeSlice :1671
	m.def("loadAdvectTimeSlice_OptInit",
	      [](PInt ID, const std::string &fname1, bool useDefoVols, bool doAligned, float partialLoadFac) {
		      const char *fn = "loadAdvectTimeSlice_OptInit";
		      int x, y, z, t = 0;
		      uniSize(fname1, x, y, z, &t);
		      if (x < 1 || y < 1 || z < 1) errMsg("Invalid src size from " + fname1);
		      debMsg(1, "Found size [" << x << "," << y << "," << z << "]," << t << " in " << fname1);
		      Lats &l = g_lats[ID.v];
		      const int keepLastT = l.lastT;  // the reference re-uses the LoadAdvectData of an ID (lastT survives, :1883-1886)
		      l.release();
		      l.lastT = keepLastT;
		      l.dd.nx = x; l.dd.ny = y; l.dd.nz = z; l.dd.nt = t;
		      l.fname = fname1;
		      l.defo = loadDefoVolume(fname1, l.dd, fn);
		      l.useDefoVols = useDefoVols;
		      if (useDefoVols) {
			      // ref :1892-1906: window of int(dimT * max(0.2, partialLoadFac)) slices, zero-initialised grids
			      const float defoVolWidth = std::max(0.2f, partialLoadFac);
			      l.Tw = (int)(t * defoVolWidth);
			      if (l.Tw < 2) errMsg("loadAdvectTimeSlice_OptInit: deformation volume too short for a defo-volume window");
			      const size_t sb = (size_t)x * y * z * 16;
			      CK(flof_malloc(ctx(), &l.vt, sb), fn);
			      CK(flof_malloc(ctx(), &l.dvt, sb * l.Tw), fn);
			      CK(flof_malloc(ctx(), &l.win[0], sb * l.Tw), fn);
			      const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
			      CK(flof_grid_set_const(ctx(), (float *)l.vt, (int64_t)x * y * z, 4, zero), fn);
			      CK(flof_grid_set_const(ctx(), (float *)l.dvt, (int64_t)x * y * z * l.Tw, 4, zero), fn);
			      CK(flof_grid_set_const(ctx(), (float *)l.win[0], (int64_t)x * y * z * l.Tw, 4, zero), fn);
			      l.vol[0] = l.defo;
			      l.numDv = 1;
			      l.doAligned = doAligned;
			      for (int i = 0; i < Lats::MAXDV; ++i) l.filepos[i] = -1;
			      debMsg(3, "Created defovol solver , " << (int)(t * 0.2) << " , defo0 " << fname1);
		      }
	      },
	      py::arg("ID"), py::arg("fname1"), py::arg("useDefoVols"), py::arg("doAligned"), py::arg("partialLoadFac") = 0.2f);
	m.def("loadAdvectTimeSlice_OptAdd", [](PInt ID, const std::string &fname) {  // ref :1914-1928
		      const char *fn = "loadAdvectTimeSlice_OptAdd";
		      auto it = g_lats.find(ID.v);
		      if (it == g_lats.end() || !it->second.defo) return;
		      Lats &l = it->second;
		      if (!l.useDefoVols) return;
		      if (l.numDv + 1 >= Lats::MAXDV) errMsg("Too many defovolumes loaded!");
		      int x, y, z, t = 0;
		      uniSize(fname, x, y, z, &t);
		      if (x != l.dd.nx || y != l.dd.ny || z != l.dd.nz || t != l.dd.nt) errMsg("grid dim doesn't match in " + fname);
		      l.vol[l.numDv] = loadDefoVolume(fname, l.dd, fn);
		      const int64_t wc = (int64_t)x * y * z * l.Tw;
		      CK(flof_malloc(ctx(), &l.win[l.numDv], (size_t)wc * 16), fn);
		      const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
		      CK(flof_grid_set_const(ctx(), (float *)l.win[l.numDv], wc, 4, zero), fn);
		      l.numDv++;
		      debMsg(3, "Added defovol " << l.numDv << " file , " << fname);
	      }, py::arg("ID"), py::arg("fname"));
	m.def("loadAdvectTimeSlice_Finish", [](PInt ID) {
		      auto it = g_lats.find(ID.v);
		      if (it == g_lats.end()) return;
		      it->second.release();
		      g_lats.erase(it);
	      }, py::arg("ID"));
	m.def("loadAdvectTimeSlice_OptRun",
	      [](PInt ID, const std::string &fname, Grid3 &dst, Grid4 &phi, float time, float blendAlpha, float loadTimeScale, const py::object &defoOffset,
	         const py::object &defoScale, const py::object &defoFactor, const py::object &overrideSize, float overrideTimeOff, const py::object &debugVel,
	         const py::object &debugVelT, bool zeroVel, float thirdAlpha, PInt bordSkip, float fourthAlpha, float defoAniFac) {
		      const char *fn = "loadAdvectTimeSlice_OptRun";
		      (void)fname; (void)debugVel; (void)debugVelT;  // (unused by the reference's optimised variant as well)
		      auto it = g_lats.find(ID.v);
		      if (it == g_lats.end() || !it->second.defo) { std::ostringstream s; s << "Load-advect data id " << ID.v << " not initialized!"; errMsg(s.str()); }
		      Lats &l = it->second;
		      requireKind(dst, K_REAL, fn, "dst"); requireKind(phi, K_REAL, fn, "phi");
		      float o[4], s[4], f[4], z[4];
		      v4arr(toV4(defoOffset), o); v4arr(toV4(defoScale), s); v4arr(toV4(defoFactor), f); v4arr(toV4(overrideSize), z);
		      if (!l.useDefoVols) {
			      if (bordSkip.v < 10) debMsg(1, "Warning - dont use for small sizes...");
			      // zeroVel: vt.setConst(0) (:2091-2094) == scaling the looked-up deformation by 0
			      CK(flof_load_advect_time_slice(ctx(), (const float *)l.defo, l.dd, dst.f(), dst.d, phi.f(), phi.d, time, blendAlpha,
			                                     loadTimeScale, o, s, f, z, overrideTimeOff, bordSkip.v, zeroVel ? 0.f : defoAniFac), fn);
			      return;
		      }
		      // ---- defo volumes, ref :2015-2089
		      if (l.numDv != 2 && l.numDv != 3) errMsg("Code currently only supports 2 deformation volumes");
		      float srcTime = 0.f, tw = 0.f, sf3[3], off3[3];
		      int t = 0, tp1 = 0;
		      flof_lats_source_time(l.dd, phi.d, time, loadTimeScale, o, s, z, &srcTime, &t, &tp1, &tw, sf3, off3);
		      debMsg(1, "Updating defo vol at " << t << " w " << tw);
		      const flof_dim4 wd = { l.dd.nx, l.dd.ny, l.dd.nz, l.Tw };
		      for (int dv = 0; dv < l.numDv; ++dv)
			      CK(flof_defovol_window_update(ctx(), (float *)l.win[dv], wd, (const float *)l.vol[dv], l.dd.nt, t, l.lastT, (float *)l.vt, &l.filepos[dv]), fn);
		      const int defovolOff = t - l.Tw / 2;
		      const float tcoord = srcTime - (float)defovolOff;
		      CK(flof_defovol_compose(ctx(), (float *)l.vt, (const float *)l.win[0], (const float *)l.win[1], l.numDv == 3 ? (const float *)l.win[2] : nullptr,
		                              (float *)l.dvt, wd, tcoord, l.doAligned ? 1 : 0, blendAlpha, thirdAlpha, fourthAlpha), fn);
		      if (zeroVel) {
			      debMsg(1, "Debug - zeroing deformation!");
			      const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
			      CK(flof_grid_set_const(ctx(), (float *)l.vt, (int64_t)l.dd.nx * l.dd.ny * l.dd.nz, 4, zero), fn);
		      }
		      l.lastT = t;
		      if (bordSkip.v < 10) debMsg(1, "Warning - dont use for small sizes...");
		      float fac[4];
		      for (int c = 0; c < 4; ++c) fac[c] = f[c] * defoAniFac;
		      const flof_dim3 vd = { l.dd.nx, l.dd.ny, l.dd.nz };
		      // blendAlpha is already accumulated in vt: the look-up runs with dt = 1 (ref :2087-2088)
		      CK(flof_lookup_slice4d_with_vel(ctx(), dst.f(), dst.d, phi.f(), phi.d, time + overrideTimeOff, 1.f, (const float *)l.vt, vd, sf3, off3, fac,
		                                      bordSkip.v), fn);
	      },
	      py::arg("ID"), py::arg("fname"), py::arg("dst"), py::arg("phi"), py::arg("time"), py::arg("blendAlpha"), py::arg("loadTimeScale"),
	      py::arg("defoOffset"), py::arg("defoScale"), py::arg("defoFactor"), py::arg("overrideSize") = py::float_(-1.), py::arg("overrideTimeOff") = 0.f,
	      py::arg("debugVel") = py::none(), py::arg("debugVelT") = py::none(), py::arg("zeroVel") = false, py::arg("thirdAlpha") = 0.f,
	      py::arg("bordSkip") = PInt{ 1 }, py::arg("fourthAlpha") = 0.f, py::arg("defoAniFac") = 1.f);
	m.def("loadAdvectTimeSlice",
	      [](PInt, const std::string &fname, Grid3 &dst, Grid4 &phi, float time, float blendAlpha, float loadTimeScale, const py::object &defoOffset,
	         const py::object &defoScale, const py::object &defoFactor, const py::object &overrideSize, float overrideTimeOff, const py::object &debugVel,
	         const py::object &debugVelT, bool zeroVel, float, PInt, float, float defoAniFac) {
		      // ref :1671-1760: the slower twin without caching (thirdAlpha / bordSkip / fourthAlpha are not supported there
		      // either).  knSemiLagrangeLookupSlice4d is KERNEL(fourd, bnd = 1) on a one-slice grid: the generated loop runs
		      // it as a 3D kernel with t = 0 over the interior cells.
		      const char *fn = "loadAdvectTimeSlice";
		      Grid3 *dv = debugVel.is_none() || py::isinstance<py::int_>(debugVel) ? nullptr : debugVel.cast<Grid3 *>();
		      Grid3 *dt = debugVelT.is_none() || py::isinstance<py::int_>(debugVelT) ? nullptr : debugVelT.cast<Grid3 *>();
		      requireKind(dst, K_REAL, fn, "dst"); requireKind(phi, K_REAL, fn, "phi");
		      if (dv) { requireKind(*dv, K_VEC3, fn, "debugVel"); if (dv->d.nx != dst.d.nx || dv->d.ny != dst.d.ny || dv->d.nz != dst.d.nz) errMsg("loadAdvectTimeSlice: debugVel size differs from dst"); }
		      if (dt) { requireKind(*dt, K_REAL, fn, "debugVelT"); if (dt->d.nx != dst.d.nx || dt->d.ny != dst.d.ny || dt->d.nz != dst.d.nz) errMsg("loadAdvectTimeSlice: debugVelT size differs from dst"); }
		      int x, y, z, t = 0;
		      uniSize(fname, x, y, z, &t);
		      if (x < 1 || y < 1 || z < 1) errMsg("Invalid src size from " + fname);
		      debMsg(1, "Found size [" << x << "," << y << "," << z << "]," << t << " in " << fname);
		      flof_dim4 dd = { x, y, z, t };
		      void *defo = loadDefoVolume(fname, dd, fn);
		      float o[4], s[4], f[4], zs[4];
		      v4arr(toV4(defoOffset), o); v4arr(toV4(defoScale), s); v4arr(toV4(defoFactor), f); v4arr(toV4(overrideSize), zs);
		      const int rc = flof_load_advect_time_slice_unopt(ctx(), (const float *)defo, dd, dst.f(), dst.d, phi.f(), phi.d, time, blendAlpha, loadTimeScale, o, s,
		                                                       f, zs, overrideTimeOff, defoAniFac, zeroVel ? 1 : 0, dv ? dv->f() : nullptr, dt ? dt->f() : nullptr);
		      flof_sync(ctx());
		      flof_free(ctx(), defo);
		      CK(rc, fn);
	      },
	      py::arg("dummyID"), py::arg("fname"), py::arg("dst"), py::arg("phi"), py::arg("time"), py::arg("blendAlpha"), py::arg("loadTimeScale"),
	      py::arg("defoOffset"), py::arg("defoScale"), py::arg("defoFactor"), py::arg("overrideSize") = py::float_(-1.), py::arg("overrideTimeOff") = 0.f,
	      py::arg("debugVel") = py::none(), py::arg("debugVelT") = py::none(), py::arg("zeroVel") = false, py::arg("thirdAlpha") = 0.f,
	      py::arg("bordSkip") = PInt{ 1 }, py::arg("fourthAlpha") = 0.f, py::arg("defoAniFac") = 1.f);

	// every plugin call / method of the reference accepts notiming=... (pclass.cpp:27-48): strip it in a thin Python shim
	py::exec(R"PY(
def _flof_wrap_notiming():
    import functools
    def wrap(f):
        @functools.wraps(f)
        def w(*a, **k):
            k.pop('notiming', None)
            return f(*a, **k)
        return w
    g = globals()
    for name, obj in list(g.items()):
        if name.startswith('_'):
            continue
        if type(obj).__name__ == 'builtin_function_or_method':
            g[name] = wrap(obj)
_flof_wrap_notiming()
del _flof_wrap_notiming
)PY", m.attr("__dict__"));
}
