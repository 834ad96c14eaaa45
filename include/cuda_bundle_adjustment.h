/*
 * cuda_bundle_adjustment.h -- the cuba::CudaBundleAdjustment interface for the B200-native engine.
 * INTERFACE DERIVED FROM fixstars/cuda-bundle-adjustment (Copyright 2020 Fixstars Corporation, Apache License 2.0,
 * http://www.apache.org/licenses/LICENSE-2.0): the virtual-method order fixes the vtable a drop-in must share.
 * Method-for-method the same abstract class as the reference's
 * include/cuda_bundle_adjustment.h:34-125 (@4390e13); create() returns the implementation in
 * libcuba_b200.so (csrc/cuba_api.cpp), which flattens the graph like the reference's
 * CudaBlockSolver::initialize (src/cuda_bundle_adjustment.cpp:115-261) and drives the C ABI of
 * include/cuba_b200.h.
 *
 * Ownership is the reference's: the caller owns vertices and edges; the optimizer mutates iP/iL,
 * the vertices' `edges` sets, and writes q/t/Xw back at the end of optimize().
 */
#ifndef CUBA_B200_CUDA_BUNDLE_ADJUSTMENT_H
#define CUBA_B200_CUDA_BUNDLE_ADJUSTMENT_H

#include "cuda_bundle_adjustment_types.h"

namespace cuba
{

class CudaBundleAdjustment
{
public:
	using Ptr = UniquePtr<CudaBundleAdjustment>;

	static Ptr create();

	virtual void addPoseVertex(PoseVertex* v) = 0;
	virtual void addLandmarkVertex(LandmarkVertex* v) = 0;
	virtual void addMonocularEdge(MonoEdge* e) = 0;
	virtual void addStereoEdge(StereoEdge* e) = 0;

	/** throws std::out_of_range for an unknown id (like the reference's map::at) */
	virtual PoseVertex* poseVertex(int id) const = 0;
	virtual LandmarkVertex* landmarkVertex(int id) const = 0;

	virtual void removePoseVertex(PoseVertex* v) = 0;
	virtual void removeLandmarkVertex(LandmarkVertex* v) = 0;
	virtual void removeEdge(BaseEdge* e) = 0;

	virtual size_t nposes() const = 0;
	virtual size_t nlandmarks() const = 0;
	virtual size_t nedges() const = 0;

	virtual void setRobustKernels(RobustKernelType kernelType, double delta, EdgeType edgeType) = 0;

	virtual void initialize() = 0;
	virtual void optimize(int niterations) = 0;
	virtual void clear() = 0;

	virtual const BatchStatistics& batchStatistics() const = 0;
	virtual const TimeProfile& timeProfile() const = 0;

	/** chi2 of one edge after optimize(); 0 for an edge whose two ends are fixed */
	virtual double chiSquared(const BaseEdge* e) const = 0;

	virtual ~CudaBundleAdjustment();
};

} // namespace cuba

#endif
