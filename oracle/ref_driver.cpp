// ref_driver.cpp -- flat-problem driver around the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
//
// Compiled together with /root/reference/src/*.{cpp,cu} into oracle/_ref/libcuba_ref.so by
// oracle/build_ref.sh.  It only uses the reference's public API (include/cuda_bundle_adjustment.h):
// builds PoseVertex / LandmarkVertex / MonoEdge / StereoEdge objects from the flat arrays, then runs
// initialize() + optimize(n) exactly like samples/sample_ba_from_file.cpp:52-57, and copies out the
// batch statistics, the estimate, the per-edge chi2 and the 8-bucket time profile.
//
// Vertex ids are the flat indices (free first, fixed last), so the reference's own index assignment
// (src/cuda_bundle_adjustment.cpp:142-200) reproduces iP = index, iL = index.
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include <cuda_runtime.h>

#include <cuda_bundle_adjustment.h>

extern "C" {

struct ref_problem {
	int Pall, numP, Lall, numL;
	const double *q, *t, *cam, *Xw;
	int E2; const int* idx2; const double* meas2; const double* omega2;
	int E3; const int* idx3; const double* meas3; const double* omega3;
};

struct Graph {
	std::vector<std::unique_ptr<cuba::PoseVertex>> P;
	std::vector<std::unique_ptr<cuba::LandmarkVertex>> L;
	std::vector<std::unique_ptr<cuba::MonoEdge>> M;
	std::vector<std::unique_ptr<cuba::StereoEdge>> S;
	cuba::CudaBundleAdjustment::Ptr opt;
};

static void build(const ref_problem* p, const int* rk_type, const double* rk_delta, Graph& g)
{
	g.opt = cuba::CudaBundleAdjustment::create();
	for (int i = 0; i < p->Pall; i++) {
		cuba::CameraParams cam;
		cam.fx = p->cam[5 * i]; cam.fy = p->cam[5 * i + 1]; cam.cx = p->cam[5 * i + 2]; cam.cy = p->cam[5 * i + 3]; cam.bf = p->cam[5 * i + 4];
		cuba::PoseVertex::Quaternion q;
		for (int k = 0; k < 4; k++) q.coeffs().data()[k] = p->q[4 * i + k];
		cuba::PoseVertex::Translation t;
		for (int k = 0; k < 3; k++) t.data()[k] = p->t[3 * i + k];
		g.P.emplace_back(new cuba::PoseVertex(i, q, t, cam, i >= p->numP));
		g.opt->addPoseVertex(g.P.back().get());
	}
	for (int i = 0; i < p->Lall; i++) {
		cuba::LandmarkVertex::Point3D X;
		for (int k = 0; k < 3; k++) X.data()[k] = p->Xw[3 * i + k];
		g.L.emplace_back(new cuba::LandmarkVertex(p->Pall + i, X, i >= p->numL));
		g.opt->addLandmarkVertex(g.L.back().get());
	}
	for (int e = 0; e < p->E2; e++) {
		cuba::MonoEdge::Measurement m;
		m.data()[0] = p->meas2[2 * e]; m.data()[1] = p->meas2[2 * e + 1];
		g.M.emplace_back(new cuba::MonoEdge(m, p->omega2[e], g.P[p->idx2[2 * e]].get(), g.L[p->idx2[2 * e + 1]].get()));
		g.opt->addMonocularEdge(g.M.back().get());
	}
	for (int e = 0; e < p->E3; e++) {
		cuba::StereoEdge::Measurement m;
		for (int k = 0; k < 3; k++) m.data()[k] = p->meas3[3 * e + k];
		g.S.emplace_back(new cuba::StereoEdge(m, p->omega3[e], g.P[p->idx3[2 * e]].get(), g.L[p->idx3[2 * e + 1]].get()));
		g.opt->addStereoEdge(g.S.back().get());
	}
	g.opt->setRobustKernels(static_cast<cuba::RobustKernelType>(rk_type[0]), rk_delta[0], cuba::EdgeType::MONOCULAR);
	g.opt->setRobustKernels(static_cast<cuba::RobustKernelType>(rk_type[1]), rk_delta[1], cuba::EdgeType::STEREO);
}

// Returns the number of batch statistics, or a negative value when no CUDA device is usable.
// seconds[0] = wall time of initialize()+optimize(niter) (the reference's timed window);
// seconds[1] = wall time of building the graph objects (not part of the reference's window).
int ref_run(const ref_problem* p, const int* rk_type, const double* rk_delta, int warmup, int niter,
	double* chi2_out, double* q, double* t, double* Xw, double* chisq_per_edge, double* seconds, double* profile8)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return -1;
	if (warmup > 0) {   // on a copy: the reference's warm-up writes its result back into the graph
		Graph w;
		build(p, rk_type, rk_delta, w);
		w.opt->initialize();
		w.opt->optimize(warmup);
	}
	const auto tb0 = std::chrono::steady_clock::now();
	Graph g;
	build(p, rk_type, rk_delta, g);
	cudaDeviceSynchronize();
	const auto t0 = std::chrono::steady_clock::now();
	g.opt->initialize();
	g.opt->optimize(niter);
	cudaDeviceSynchronize();
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) {
		seconds[0] = std::chrono::duration<double>(t1 - t0).count();
		seconds[1] = std::chrono::duration<double>(t0 - tb0).count();
	}
	const auto& stats = g.opt->batchStatistics();
	for (size_t i = 0; i < stats.size(); i++) if (chi2_out) chi2_out[i] = stats[i].chi2;
	for (int i = 0; i < p->Pall; i++) {
		if (q) for (int k = 0; k < 4; k++) q[4 * i + k] = g.P[i]->q.coeffs().data()[k];
		if (t) for (int k = 0; k < 3; k++) t[3 * i + k] = g.P[i]->t.data()[k];
	}
	if (Xw) for (int i = 0; i < p->Lall; i++) for (int k = 0; k < 3; k++) Xw[3 * i + k] = g.L[i]->Xw.data()[k];
	if (chisq_per_edge) {
		for (int e = 0; e < p->E2; e++) chisq_per_edge[e] = g.opt->chiSquared(g.M[e].get());
		for (int e = 0; e < p->E3; e++) chisq_per_edge[p->E2 + e] = g.opt->chiSquared(g.S[e].get());
	}
	if (profile8) {
		int k = 0;
		for (const auto& kv : g.opt->timeProfile()) { if (k < 8) profile8[k++] = kv.second; }
	}
	return static_cast<int>(stats.size());
}

}  // extern "C"
