// cuba_api.cpp -- cuba::CudaBundleAdjustment on top of the C ABI (include/cuba_b200.h).
//
// Host-side mirror of the reference's graph container + initialize():
//   graph container      reference src/cuda_bundle_adjustment.cpp:677-781
//   index assignment     cpp:142-200  (ascending id, free first, fixed appended, edge-less vertices skipped)
//   edge flattening      cpp:202-243  (monocular ids first, then stereo; both-fixed edges dropped)
//   finalize/getChiSqs   cpp:512-543  (write-back of q,t,Xw into the caller's vertices, per-edge chi2)
// Differences, on purpose:
//   * edges are kept in insertion order (the reference iterates an unordered_set, so its order depends on heap addresses);
//   * the per-pose cameras are rebuilt on every initialize() (the reference never clears cameras_);
//   * removing a vertex copies its edge set before erasing from it;
//   * initialize() is built for repeated local-BA calls: vertices are looked up through hash maps and walked in a cached
//     ascending-id order (re-sorted only after an add/remove), removals are O(1) tombstones compacted at the next initialize(),
//     the flat arrays live in page-locked memory (the engine's H2D copies are plain DMA) and the edge pass is split over a few
//     host threads; the pointer -> chi2 map the reference rebuilds in every optimize() (cpp:541-542) is built on the first
//     chiSquared() call instead.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cuba_b200.h"
#include "../../include/cuda_bundle_adjustment.h"

namespace cuba
{

namespace
{

// grow-only host array in page-locked memory (pageable when no CUDA device is usable: initialize() itself needs no GPU)
template <typename T>
class HostBuf
{
public:
	HostBuf() {}
	HostBuf(const HostBuf&) = delete;
	HostBuf& operator=(const HostBuf&) = delete;
	~HostBuf() { release(); }
	void resize(size_t n)
	{
		if (n > cap_) {
			release();
			cap_ = n + n / 8 + 16;
			void* q = nullptr;
			if (cudaMallocHost(&q, cap_ * sizeof(T)) == cudaSuccess) { p_ = static_cast<T*>(q); pinned_ = true; }
			else { cudaGetLastError(); p_ = static_cast<T*>(std::malloc(cap_ * sizeof(T))); pinned_ = false; if (!p_) throw std::bad_alloc(); }
		}
		n_ = n;
	}
	T* data() { return p_; }
	const T* data() const { return p_; }
	size_t size() const { return n_; }
	T& operator[](size_t i) { return p_[i]; }
	const T& operator[](size_t i) const { return p_[i]; }
private:
	void release()
	{
		if (p_) { if (pinned_) cudaFreeHost(p_); else std::free(p_); }
		p_ = nullptr; cap_ = 0; n_ = 0;
	}
	T* p_ = nullptr; size_t n_ = 0, cap_ = 0; bool pinned_ = false;
};

// Host loops over vertex / edge objects are pointer chasing (memory-latency bound): a few threads, each prefetching ahead.
static unsigned host_threads(size_t n, size_t grain)
{
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	return static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(std::min<size_t>(16, hw), n / grain + 1)));
}

// runs fn(t, begin, end) for the t-th of nthreads equal slices of [0, n)
template <class F>
static void parallel_slices(size_t n, unsigned nthreads, F fn)
{
	if (nthreads <= 1) { fn(0u, size_t(0), n); return; }
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(fn, t, n * t / nthreads, n * (t + 1) / nthreads);
	fn(0u, size_t(0), n / nthreads);
	for (auto& th : pool) th.join();
}

template <class F>
static void parallel_for(size_t n, F fn)
{
	parallel_slices(n, host_threads(n, 32768), [&fn](unsigned, size_t b, size_t e) { fn(b, e); });
}

class Impl : public CudaBundleAdjustment
{
public:
	Impl()
	{
		kernels_[0] = { 0, 0.0 };
		kernels_[1] = { 0, 0.0 };
	}
	~Impl() override { if (engine_) cuba_engine_destroy(engine_); }

	void addPoseVertex(PoseVertex* v) override { if (poses_.insert({ v->id, v }).second) orderDirty_ = true; }
	void addLandmarkVertex(LandmarkVertex* v) override { if (landmarks_.insert({ v->id, v }).second) orderDirty_ = true; }

	void addMonocularEdge(MonoEdge* e) override
	{
		if (!member_.insert(e).second) return;
		edgeVersion_++;
		mono_.push_back(e);
		e->vertexP->edges.insert(e); e->vertexL->edges.insert(e);
	}
	void addStereoEdge(StereoEdge* e) override
	{
		if (!member_.insert(e).second) return;
		edgeVersion_++;
		stereo_.push_back(e);
		e->vertexP->edges.insert(e); e->vertexL->edges.insert(e);
	}

	PoseVertex* poseVertex(int id) const override { return poses_.at(id); }
	LandmarkVertex* landmarkVertex(int id) const override { return landmarks_.at(id); }

	void removePoseVertex(PoseVertex* v) override
	{
		auto it = poses_.find(v->id);
		if (it == poses_.end()) return;
		const std::vector<BaseEdge*> es(it->second->edges.begin(), it->second->edges.end());
		for (auto e : es) removeEdge(e);
		poses_.erase(it);
		orderDirty_ = true;
	}
	void removeLandmarkVertex(LandmarkVertex* v) override
	{
		auto it = landmarks_.find(v->id);
		if (it == landmarks_.end()) return;
		const std::vector<BaseEdge*> es(it->second->edges.begin(), it->second->edges.end());
		for (auto e : es) removeEdge(e);
		landmarks_.erase(it);
		orderDirty_ = true;
	}
	// O(1): the slot in the insertion-ordered list becomes a tombstone, dropped by the next initialize()
	void removeEdge(BaseEdge* e) override
	{
		if (auto p = e->poseVertex()) p->edges.erase(e);
		if (auto l = e->landmarkVertex()) l->edges.erase(e);
		if (!member_.erase(e)) return;
		edgeVersion_++;
		tombstones_[e]++;
		(e->dim() == 2 ? deadMono_ : deadStereo_)++;
	}

	size_t nposes() const override { return poses_.size(); }
	size_t nlandmarks() const override { return landmarks_.size(); }
	size_t nedges() const override { return mono_.size() + stereo_.size() - deadMono_ - deadStereo_; }

	void setRobustKernels(RobustKernelType kernelType, double delta, EdgeType edgeType) override
	{
		kernels_[static_cast<int>(edgeType)] = { static_cast<int>(kernelType), delta };
	}

	void initialize() override
	{
		const auto t0 = std::chrono::steady_clock::now();
		compactEdges();
		if (orderDirty_) {
			orderP_.clear(); orderL_.clear();
			orderP_.reserve(poses_.size()); orderL_.reserve(landmarks_.size());
			for (const auto& kv : poses_) orderP_.push_back(kv.second);
			for (const auto& kv : landmarks_) orderL_.push_back(kv.second);
			std::sort(orderP_.begin(), orderP_.end(), [](const PoseVertex* a, const PoseVertex* b) { return a->id < b->id; });
			std::sort(orderL_.begin(), orderL_.end(), [](const LandmarkVertex* a, const LandmarkVertex* b) { return a->id < b->id; });
			orderDirty_ = false;
		}
		// index assignment: ascending id, free vertices first, fixed ones appended, vertices without edges skipped
		// (if neither the numbering nor the edge lists moved since the last initialize(), the edge pass below only refreshes values)
		prevP_.swap(vP_);
		const int prevNumP = numP_, prevNumL = numL_;
		vP_.clear();
		vP_.reserve(orderP_.size());
		size_t nFixedP = 0, nFixedL = 0;
		for (PoseVertex* v : orderP_) { if (v->edges.empty()) continue; if (v->fixed) nFixedP++; else { v->iP = static_cast<int>(vP_.size()); vP_.push_back(v); } }
		numP_ = static_cast<int>(vP_.size());
		if (nFixedP) for (PoseVertex* v : orderP_) if (v->fixed && !v->edges.empty()) { v->iP = static_cast<int>(vP_.size()); vP_.push_back(v); }
		// landmarks (many): pass 1 classifies every vertex (0 = no edges, 1 = free, 2 = fixed) and counts per slice, pass 2 writes
		// index, list entry and coordinates at the slice's offsets -- same numbering as a serial walk in ascending id
		{
			const size_t nl = orderL_.size();
			const unsigned nt = host_threads(nl, 16384);
			cls_.resize(nl);
			std::vector<size_t> cntFree(nt + 1, 0), cntFixed(nt + 1, 0);
			parallel_slices(nl, nt, [&](unsigned tix, size_t b, size_t e) {
				size_t nf = 0, nx = 0;
				for (size_t i = b; i < e; i++) {
					if (i + 8 < e) __builtin_prefetch(orderL_[i + 8]);
					const LandmarkVertex* v = orderL_[i];
					const unsigned char c = v->edges.empty() ? 0 : (v->fixed ? 2 : 1);
					cls_[i] = c; nf += c == 1; nx += c == 2;
				}
				cntFree[tix + 1] = nf; cntFixed[tix + 1] = nx;
			});
			for (unsigned t = 0; t < nt; t++) { cntFree[t + 1] += cntFree[t]; cntFixed[t + 1] += cntFixed[t]; }
			numL_ = static_cast<int>(cntFree[nt]);
			nFixedL = cntFixed[nt];
			sameNumbering_ = vL_.size() == cntFree[nt] + cntFixed[nt] && numL_ == prevNumL;
			vL_.resize(cntFree[nt] + cntFixed[nt]);
			Xw_.resize(3 * vL_.size());
			std::vector<unsigned char> moved(nt, 0);
			parallel_slices(nl, nt, [&](unsigned tix, size_t b, size_t e) {
				size_t wf = cntFree[tix], wx = cntFree[nt] + cntFixed[tix];
				unsigned char mv = 0;
				for (size_t i = b; i < e; i++) {
					if (i + 8 < e) __builtin_prefetch(orderL_[i + 8], 1);
					if (!cls_[i]) continue;
					LandmarkVertex* v = orderL_[i];
					const size_t w = cls_[i] == 1 ? wf++ : wx++;
					mv |= vL_[w] != v;
					v->iL = static_cast<int>(w); vL_[w] = v;
					for (int k = 0; k < 3; k++) Xw_[3 * w + k] = v->Xw.data()[k];
				}
				moved[tix] = mv;
			});
			for (unsigned char mv : moved) if (mv) sameNumbering_ = false;
		}
		sameNumbering_ = sameNumbering_ && numP_ == prevNumP && vP_ == prevP_;

		q_.resize(4 * vP_.size()); t_.resize(3 * vP_.size()); cam_.resize(5 * vP_.size());
		for (size_t i = 0; i < vP_.size(); i++) {
			const PoseVertex* v = vP_[i];
			for (int k = 0; k < 4; k++) q_[4 * i + k] = v->q.coeffs().data()[k];
			for (int k = 0; k < 3; k++) t_[3 * i + k] = v->t.data()[k];
			cam_[5 * i] = v->camera.fx; cam_[5 * i + 1] = v->camera.fy; cam_[5 * i + 2] = v->camera.cx;
			cam_[5 * i + 3] = v->camera.cy; cam_[5 * i + 4] = v->camera.bf;
		}

		// edges: every list entry is written at its own position by a few threads; only if an edge with both ends fixed
		// turned up (they are dropped, cpp:210-211) the arrays are closed up afterwards
		idx2_.resize(2 * mono_.size()); meas2_.resize(2 * mono_.size()); om2_.resize(mono_.size());
		idx3_.resize(2 * stereo_.size()); meas3_.resize(3 * stereo_.size()); om3_.resize(stereo_.size());
		const size_t n2 = mono_.size(), n3 = stereo_.size(), total = n2 + n3;
		const unsigned nthreads = host_threads(total, 65536);
		if (flatValid_ && sameNumbering_ && flatVersion_ == edgeVersion_ && om2_.size() == n2 && om3_.size() == n3) {
			// same edges in the same order between the same indices, none dropped: the index pairs and the position -> edge lists of
			// the last initialize() still hold; the caller may have edited measurements / information in place
			parallel_slices(total, nthreads, [&](unsigned, size_t b, size_t e) {
				for (size_t k = b; k < e; k++) {
					if (k + 16 < e) __builtin_prefetch(k + 16 < n2 ? static_cast<const void*>(mono_[k + 16]) : static_cast<const void*>(stereo_[k + 16 - n2]));
					if (k < n2) {
						const MonoEdge* ed = mono_[k];
						meas2_[2 * k] = ed->measurement.data()[0]; meas2_[2 * k + 1] = ed->measurement.data()[1];
						om2_[k] = ed->information;
					} else {
						const size_t j = k - n2;
						const StereoEdge* ed = stereo_[j];
						for (int c = 0; c < 3; c++) meas3_[3 * j + c] = ed->measurement.data()[c];
						om3_[j] = ed->information;
					}
				}
			});
			stats_.clear();
			uploaded_ = false;
			initSeconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			initialized_ = true;
			return;
		}
		auto am = takeList(mono_.size());
		auto as = takeList(stereo_.size());
		std::vector<size_t> dropped(nthreads, 0);
		parallel_slices(total, nthreads, [&](unsigned tix, size_t b, size_t e) {
			// two prefetch distances: the edge object 16 ahead, its landmark (read through the edge) 8 ahead
			auto edgeAt = [&](size_t k) -> const BaseEdge* { return k < n2 ? static_cast<const BaseEdge*>(mono_[k]) : static_cast<const BaseEdge*>(stereo_[k - n2]); };
			size_t drop = 0;
			for (size_t k = b; k < e; k++) {
				if (k + 16 < e) __builtin_prefetch(edgeAt(k + 16));
				if (k + 8 < e) __builtin_prefetch(k + 8 < n2 ? static_cast<const void*>(mono_[k + 8]->vertexL) : static_cast<const void*>(stereo_[k + 8 - n2]->vertexL));
				if (k < n2) {
					const MonoEdge* ed = mono_[k];
					const PoseVertex* vp = ed->vertexP; const LandmarkVertex* vl = ed->vertexL;
					if (vp->fixed && vl->fixed) drop++;
					(*am)[k] = ed;
					idx2_[2 * k] = vp->iP; idx2_[2 * k + 1] = vl->iL;
					meas2_[2 * k] = ed->measurement.data()[0]; meas2_[2 * k + 1] = ed->measurement.data()[1];
					om2_[k] = ed->information;
				} else {
					const size_t j = k - n2;
					const StereoEdge* ed = stereo_[j];
					const PoseVertex* vp = ed->vertexP; const LandmarkVertex* vl = ed->vertexL;
					if (vp->fixed && vl->fixed) drop++;
					(*as)[j] = ed;
					idx3_[2 * j] = vp->iP; idx3_[2 * j + 1] = vl->iL;
					for (int c = 0; c < 3; c++) meas3_[3 * j + c] = ed->measurement.data()[c];
					om3_[j] = ed->information;
				}
			}
			dropped[tix] = drop;
		});
		size_t ndrop = 0;
		for (size_t d : dropped) ndrop += d;
		if (ndrop) {
			size_t w = 0;
			for (size_t k = 0; k < n2; k++) {
				const MonoEdge* ed = mono_[k];
				if (ed->vertexP->fixed && ed->vertexL->fixed) continue;
				(*am)[w] = ed; idx2_[2 * w] = idx2_[2 * k]; idx2_[2 * w + 1] = idx2_[2 * k + 1];
				meas2_[2 * w] = meas2_[2 * k]; meas2_[2 * w + 1] = meas2_[2 * k + 1]; om2_[w] = om2_[k]; w++;
			}
			am->resize(w); idx2_.resize(2 * w); meas2_.resize(2 * w); om2_.resize(w);
			w = 0;
			for (size_t j = 0; j < n3; j++) {
				const StereoEdge* ed = stereo_[j];
				if (ed->vertexP->fixed && ed->vertexL->fixed) continue;
				(*as)[w] = ed; idx3_[2 * w] = idx3_[2 * j]; idx3_[2 * w + 1] = idx3_[2 * j + 1];
				for (int c = 0; c < 3; c++) meas3_[3 * w + c] = meas3_[3 * j + c];
				om3_[w] = om3_[j]; w++;
			}
			as->resize(w); idx3_.resize(2 * w); meas3_.resize(3 * w); om3_.resize(w);
		}
		activeMono_ = am; activeStereo_ = as;
		flatValid_ = ndrop == 0; flatVersion_ = edgeVersion_;
		stats_.clear();
		uploaded_ = false;
		initSeconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		initialized_ = true;
	}

	void optimize(int niterations) override
	{
		if (!initialized_) initialize();
		ensureEngine();
		check(cuba_engine_set_robust_kernel(engine_, CUBA_EDGE_MONOCULAR, kernels_[0].type, kernels_[0].delta));
		check(cuba_engine_set_robust_kernel(engine_, CUBA_EDGE_STEREO, kernels_[1].type, kernels_[1].delta));
		if (!uploaded_) {
			// the reference builds its structure inside the first optimize() iteration (cpp:804-805)
			cuba_problem p;
			flatProblem(p);
			check(cuba_engine_set_problem(engine_, &p));
			uploaded_ = true;
		}
		std::vector<cuba_iter_stat> st(niterations > 0 ? niterations : 1);
		int n = 0;
		if (niterations > 0) check(cuba_engine_optimize(engine_, niterations, st.data(), &n));
		// the reference appends to stats_ across optimize() calls (cleared by initialize()), cpp:848
		for (int i = 0; i < n; i++) stats_.push_back({ st[i].iteration, st[i].chi2 });

		// finalize(): write the estimate back into the caller's vertices (fixed ones too)
		check(cuba_engine_get_state(engine_, q_.data(), t_.data(), Xw_.data()));
		for (size_t i = 0; i < vP_.size(); i++) {
			for (int k = 0; k < 4; k++) vP_[i]->q.coeffs().data()[k] = q_[4 * i + k];
			for (int k = 0; k < 3; k++) vP_[i]->t.data()[k] = t_[3 * i + k];
		}
		parallel_for(vL_.size(), [this](size_t b, size_t e) { for (size_t i = b; i < e; i++) for (int k = 0; k < 3; k++) vL_[i]->Xw.data()[k] = Xw_[3 * i + k]; });

		// getChiSqs(): the values now, the pointer -> value map when somebody asks (chiSquared)
		chi_.resize(om2_.size() + om3_.size());
		if (chi_.size()) check(cuba_engine_get_chi2(engine_, chi_.data()));
		chiMono_ = activeMono_; chiStereo_ = activeStereo_;
		chiIndexValid_ = false;

		double sec[CUBA_PROF_NUM] = { 0 };
		check(cuba_engine_get_profile(engine_, sec));
		static const char* names[CUBA_PROF_NUM] = { "0: Initialize Optimizer", "1: Build Structure", "2: Compute Error",
			"3: Build System", "4: Schur Complement", "5: Symbolic Decomposition", "6: Numerical Decomposition", "7: Update Solution" };
		profile_.clear();
		for (int i = 0; i < CUBA_PROF_NUM; i++) profile_[names[i]] = sec[i];
		profile_[names[0]] += initSeconds_;
	}

	// the flat arrays of the last initialize() (also behind cuba_debug_dropin_problem)
	bool flatProblem(cuba_problem& p) const
	{
		if (!initialized_) return false;
		p.Pall = static_cast<int32_t>(vP_.size()); p.numP = numP_; p.Lall = static_cast<int32_t>(vL_.size()); p.numL = numL_;
		p.q = q_.data(); p.t = t_.data(); p.cam = cam_.data(); p.Xw = Xw_.data();
		p.E2 = static_cast<int32_t>(om2_.size()); p.idx2 = idx2_.data(); p.meas2 = meas2_.data(); p.omega2 = om2_.data();
		p.E3 = static_cast<int32_t>(om3_.size()); p.idx3 = idx3_.data(); p.meas3 = meas3_.data(); p.omega3 = om3_.data();
		return true;
	}

	void clear() override
	{
		poses_.clear(); landmarks_.clear(); mono_.clear(); stereo_.clear(); member_.clear(); tombstones_.clear();
		deadMono_ = deadStereo_ = 0; orderDirty_ = true;
		edgeVersion_++; flatValid_ = false;
		stats_.clear(); initialized_ = false; uploaded_ = false;
	}

	const BatchStatistics& batchStatistics() const override { return stats_; }
	const TimeProfile& timeProfile() const override { return profile_; }
	double chiSquared(const BaseEdge* e) const override
	{
		if (!chiIndexValid_) {
			chiIndex_.clear();
			const size_t n2 = chiMono_ ? chiMono_->size() : 0, n3 = chiStereo_ ? chiStereo_->size() : 0;
			chiIndex_.reserve(n2 + n3);
			for (size_t i = 0; i < n2; i++) chiIndex_[(*chiMono_)[i]] = i;
			for (size_t i = 0; i < n3; i++) chiIndex_[(*chiStereo_)[i]] = n2 + i;
			chiIndexValid_ = true;
		}
		auto it = chiIndex_.find(e);
		return it == chiIndex_.end() || it->second >= chi_.size() ? 0.0 : chi_[it->second];
	}

private:
	struct Kernel { int type; double delta; };

	// drops the tombstoned slots (earliest occurrences first: an edge removed and added again keeps its new position)
	void compactEdges()
	{
		if (tombstones_.empty()) return;
		auto sweep = [this](auto& vec) {
			size_t w = 0;
			for (size_t k = 0; k < vec.size(); k++) {
				auto it = tombstones_.find(vec[k]);
				if (it != tombstones_.end()) { if (--it->second == 0) tombstones_.erase(it); continue; }
				vec[w++] = vec[k];
			}
			vec.resize(w);
		};
		sweep(mono_); sweep(stereo_);
		tombstones_.clear();
		deadMono_ = deadStereo_ = 0;
	}

	// a position -> edge list for chiSquared(): the previous optimize()'s lists stay alive in chiMono_/chiStereo_, so the
	// lists rotate through a small pool instead of being allocated (and page-faulted in) on every initialize()
	std::shared_ptr<std::vector<const BaseEdge*>> takeList(size_t n)
	{
		for (auto& l : listPool_)
			if (l.use_count() == 1) { l->resize(n); return l; }
		listPool_.push_back(std::make_shared<std::vector<const BaseEdge*>>(n));
		return listPool_.back();
	}

	void ensureEngine()
	{
		if (engine_) return;
		cuba_config cfg{};
		cfg.device = -1; cfg.deterministic = 1;
#ifdef USE_FLOAT32
		cfg.use_fp32 = 1;
#endif
		check(cuba_engine_create(&cfg, &engine_));
	}
	static void check(int rc)
	{
		// the reference prints CUDA errors and continues (src/macro.h:22-27); failing loudly is safer
		if (rc != CUBA_OK) throw std::runtime_error(std::string("cuba_b200: ") + cuba_last_error());
	}

	std::unordered_map<int, PoseVertex*> poses_;
	std::unordered_map<int, LandmarkVertex*> landmarks_;
	std::vector<PoseVertex*> orderP_;          // ascending id, valid while !orderDirty_
	std::vector<LandmarkVertex*> orderL_;
	bool orderDirty_ = true;
	std::vector<MonoEdge*> mono_;              // insertion order, may hold tombstoned slots until the next initialize()
	std::vector<StereoEdge*> stereo_;
	std::unordered_set<const BaseEdge*> member_;
	std::unordered_map<const BaseEdge*, int> tombstones_;
	size_t deadMono_ = 0, deadStereo_ = 0;
	Kernel kernels_[2];

	std::vector<PoseVertex*> vP_, prevP_;
	std::vector<LandmarkVertex*> vL_;
	size_t edgeVersion_ = 0, flatVersion_ = 0;   // edits of the edge lists / the version the flat edge arrays were built from
	bool flatValid_ = false, sameNumbering_ = false;
	std::shared_ptr<const std::vector<const BaseEdge*>> activeMono_, activeStereo_, chiMono_, chiStereo_;
	std::vector<std::shared_ptr<std::vector<const BaseEdge*>>> listPool_;
	std::vector<unsigned char> cls_;
	int numP_ = 0, numL_ = 0;
	HostBuf<double> q_, t_, cam_, Xw_, meas2_, om2_, meas3_, om3_, chi_;
	HostBuf<int32_t> idx2_, idx3_;
	bool initialized_ = false, uploaded_ = false;
	double initSeconds_ = 0;

	cuba_engine* engine_ = nullptr;
	BatchStatistics stats_;
	TimeProfile profile_;
	mutable std::unordered_map<const BaseEdge*, size_t> chiIndex_;
	mutable bool chiIndexValid_ = false;
};

} // namespace

CudaBundleAdjustment::Ptr CudaBundleAdjustment::create() { return Ptr(new Impl()); }

// include/cuba_b200.h: cuba_debug_dropin_problem
static bool dropin_problem(CudaBundleAdjustment* obj, cuba_problem* out)
{
	const Impl* impl = dynamic_cast<const Impl*>(obj);
	return impl && out && impl->flatProblem(*out);
}
CudaBundleAdjustment::~CudaBundleAdjustment() {}

} // namespace cuba

extern "C" int cuba_debug_dropin_problem(void* dropin, cuba_problem* out)
{
	return cuba::dropin_problem(static_cast<cuba::CudaBundleAdjustment*>(dropin), out) ? CUBA_OK : CUBA_ERR_STATE;
}
