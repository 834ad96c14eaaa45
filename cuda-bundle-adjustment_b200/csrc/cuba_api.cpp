// cuba_api.cpp -- cuba::CudaBundleAdjustment on top of the C ABI (include/cuba_b200.h).
//
// Host-side mirror of the reference's graph container + initialize():
//   graph container      reference src/cuda_bundle_adjustment.cpp:677-781
//   index assignment     cpp:142-200  (ascending id, free first, fixed appended, edge-less vertices skipped)
//   edge flattening      cpp:202-243  (monocular ids first, then stereo; both-fixed edges dropped)
//   finalize/getChiSqs   cpp:512-543  (write-back of q,t,Xw into the caller's vertices, per-edge chi2)
// Differences, on purpose: edges are kept in insertion order (the reference iterates an unordered_set, so
// its order depends on heap addresses); the per-pose cameras are rebuilt on every initialize() (the
// reference never clears cameras_); removing a vertex copies its edge set before erasing from it.
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cuba_b200.h"
#include "../../include/cuda_bundle_adjustment.h"

namespace cuba
{

namespace
{

class Impl : public CudaBundleAdjustment
{
public:
	Impl()
	{
		kernels_[0] = { 0, 0.0 };
		kernels_[1] = { 0, 0.0 };
	}
	~Impl() override { if (engine_) cuba_engine_destroy(engine_); }

	void addPoseVertex(PoseVertex* v) override { poses_.insert({ v->id, v }); }
	void addLandmarkVertex(LandmarkVertex* v) override { landmarks_.insert({ v->id, v }); }

	void addMonocularEdge(MonoEdge* e) override
	{
		if (monoPos_.count(e)) return;
		monoPos_[e] = mono_.size(); mono_.push_back(e);
		e->vertexP->edges.insert(e); e->vertexL->edges.insert(e);
	}
	void addStereoEdge(StereoEdge* e) override
	{
		if (stereoPos_.count(e)) return;
		stereoPos_[e] = stereo_.size(); stereo_.push_back(e);
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
	}
	void removeLandmarkVertex(LandmarkVertex* v) override
	{
		auto it = landmarks_.find(v->id);
		if (it == landmarks_.end()) return;
		const std::vector<BaseEdge*> es(it->second->edges.begin(), it->second->edges.end());
		for (auto e : es) removeEdge(e);
		landmarks_.erase(it);
	}
	void removeEdge(BaseEdge* e) override
	{
		if (auto p = e->poseVertex()) p->edges.erase(e);
		if (auto l = e->landmarkVertex()) l->edges.erase(e);
		if (e->dim() == 2) eraseFrom(mono_, monoPos_, static_cast<MonoEdge*>(e));
		if (e->dim() == 3) eraseFrom(stereo_, stereoPos_, static_cast<StereoEdge*>(e));
	}

	size_t nposes() const override { return poses_.size(); }
	size_t nlandmarks() const override { return landmarks_.size(); }
	size_t nedges() const override { return mono_.size() + stereo_.size(); }

	void setRobustKernels(RobustKernelType kernelType, double delta, EdgeType edgeType) override
	{
		kernels_[static_cast<int>(edgeType)] = { static_cast<int>(kernelType), delta };
	}

	void initialize() override
	{
		const auto t0 = std::chrono::steady_clock::now();
		vP_.clear(); vL_.clear(); activeMono_.clear(); activeStereo_.clear();
		std::vector<PoseVertex*> fixedP; std::vector<LandmarkVertex*> fixedL;
		for (const auto& kv : poses_) {
			PoseVertex* v = kv.second;
			if (v->edges.empty()) continue;
			if (!v->fixed) { v->iP = static_cast<int>(vP_.size()); vP_.push_back(v); } else fixedP.push_back(v);
		}
		for (const auto& kv : landmarks_) {
			LandmarkVertex* v = kv.second;
			if (v->edges.empty()) continue;
			if (!v->fixed) { v->iL = static_cast<int>(vL_.size()); vL_.push_back(v); } else fixedL.push_back(v);
		}
		numP_ = static_cast<int>(vP_.size()); numL_ = static_cast<int>(vL_.size());
		for (auto v : fixedP) { v->iP = static_cast<int>(vP_.size()); vP_.push_back(v); }
		for (auto v : fixedL) { v->iL = static_cast<int>(vL_.size()); vL_.push_back(v); }

		q_.resize(4 * vP_.size()); t_.resize(3 * vP_.size()); cam_.resize(5 * vP_.size()); Xw_.resize(3 * vL_.size());
		for (size_t i = 0; i < vP_.size(); i++) {
			const PoseVertex* v = vP_[i];
			for (int k = 0; k < 4; k++) q_[4 * i + k] = v->q.coeffs().data()[k];
			for (int k = 0; k < 3; k++) t_[3 * i + k] = v->t.data()[k];
			cam_[5 * i] = v->camera.fx; cam_[5 * i + 1] = v->camera.fy; cam_[5 * i + 2] = v->camera.cx;
			cam_[5 * i + 3] = v->camera.cy; cam_[5 * i + 4] = v->camera.bf;
		}
		for (size_t i = 0; i < vL_.size(); i++) for (int k = 0; k < 3; k++) Xw_[3 * i + k] = vL_[i]->Xw.data()[k];

		idx2_.clear(); meas2_.clear(); om2_.clear(); idx3_.clear(); meas3_.clear(); om3_.clear();
		for (auto e : mono_) {
			if (e->vertexP->fixed && e->vertexL->fixed) continue;
			activeMono_.push_back(e);
			idx2_.push_back(e->vertexP->iP); idx2_.push_back(e->vertexL->iL);
			meas2_.push_back(e->measurement.data()[0]); meas2_.push_back(e->measurement.data()[1]);
			om2_.push_back(e->information);
		}
		for (auto e : stereo_) {
			if (e->vertexP->fixed && e->vertexL->fixed) continue;
			activeStereo_.push_back(e);
			idx3_.push_back(e->vertexP->iP); idx3_.push_back(e->vertexL->iL);
			for (int k = 0; k < 3; k++) meas3_.push_back(e->measurement.data()[k]);
			om3_.push_back(e->information);
		}
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
			p.Pall = static_cast<int32_t>(vP_.size()); p.numP = numP_; p.Lall = static_cast<int32_t>(vL_.size()); p.numL = numL_;
			p.q = q_.data(); p.t = t_.data(); p.cam = cam_.data(); p.Xw = Xw_.data();
			p.E2 = static_cast<int32_t>(om2_.size()); p.idx2 = idx2_.data(); p.meas2 = meas2_.data(); p.omega2 = om2_.data();
			p.E3 = static_cast<int32_t>(om3_.size()); p.idx3 = idx3_.data(); p.meas3 = meas3_.data(); p.omega3 = om3_.data();
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
		for (size_t i = 0; i < vL_.size(); i++) for (int k = 0; k < 3; k++) vL_[i]->Xw.data()[k] = Xw_[3 * i + k];

		// getChiSqs()
		chi_.assign(om2_.size() + om3_.size(), 0.0);
		if (!chi_.empty()) check(cuba_engine_get_chi2(engine_, chi_.data()));
		chiOf_.clear();
		for (size_t i = 0; i < activeMono_.size(); i++) chiOf_[activeMono_[i]] = chi_[i];
		for (size_t i = 0; i < activeStereo_.size(); i++) chiOf_[activeStereo_[i]] = chi_[activeMono_.size() + i];

		double sec[CUBA_PROF_NUM] = { 0 };
		check(cuba_engine_get_profile(engine_, sec));
		static const char* names[CUBA_PROF_NUM] = { "0: Initialize Optimizer", "1: Build Structure", "2: Compute Error",
			"3: Build System", "4: Schur Complement", "5: Symbolic Decomposition", "6: Numerical Decomposition", "7: Update Solution" };
		profile_.clear();
		for (int i = 0; i < CUBA_PROF_NUM; i++) profile_[names[i]] = sec[i];
		profile_[names[0]] += initSeconds_;
	}

	void clear() override
	{
		poses_.clear(); landmarks_.clear(); mono_.clear(); stereo_.clear(); monoPos_.clear(); stereoPos_.clear();
		stats_.clear(); initialized_ = false; uploaded_ = false;
	}

	const BatchStatistics& batchStatistics() const override { return stats_; }
	const TimeProfile& timeProfile() const override { return profile_; }
	double chiSquared(const BaseEdge* e) const override
	{
		auto it = chiOf_.find(e);
		return it == chiOf_.end() ? 0.0 : it->second;
	}

private:
	struct Kernel { int type; double delta; };

	template <class E>
	static void eraseFrom(std::vector<E*>& vec, std::unordered_map<const E*, size_t>& pos, E* e)
	{
		auto it = pos.find(e);
		if (it == pos.end()) return;
		const size_t i = it->second;
		pos.erase(it);
		vec.erase(vec.begin() + i);               // keeps insertion order
		for (size_t k = i; k < vec.size(); k++) pos[vec[k]] = k;
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

	std::map<int, PoseVertex*> poses_;
	std::map<int, LandmarkVertex*> landmarks_;
	std::vector<MonoEdge*> mono_;
	std::vector<StereoEdge*> stereo_;
	std::unordered_map<const MonoEdge*, size_t> monoPos_;
	std::unordered_map<const StereoEdge*, size_t> stereoPos_;
	Kernel kernels_[2];

	std::vector<PoseVertex*> vP_;
	std::vector<LandmarkVertex*> vL_;
	std::vector<MonoEdge*> activeMono_;
	std::vector<StereoEdge*> activeStereo_;
	int numP_ = 0, numL_ = 0;
	std::vector<double> q_, t_, cam_, Xw_, meas2_, om2_, meas3_, om3_, chi_;
	std::vector<int32_t> idx2_, idx3_;
	bool initialized_ = false, uploaded_ = false;
	double initSeconds_ = 0;

	cuba_engine* engine_ = nullptr;
	BatchStatistics stats_;
	TimeProfile profile_;
	std::unordered_map<const BaseEdge*, double> chiOf_;
};

} // namespace

CudaBundleAdjustment::Ptr CudaBundleAdjustment::create() { return Ptr(new Impl()); }
CudaBundleAdjustment::~CudaBundleAdjustment() {}

} // namespace cuba
