/*
 * cuda_bundle_adjustment_types.h -- graph types of the cuba:: API for the B200-native engine.
 *
 * INTERFACE DECLARATIONS DERIVED FROM fixstars/cuda-bundle-adjustment (include/cuda_bundle_adjustment_types.h @4390e13,
 * Copyright 2020 Fixstars Corporation, Apache License 2.0, http://www.apache.org/licenses/LICENSE-2.0): a source-compatible
 * drop-in must reproduce the type names, the member order and the constructor signatures of Edge<DIM>, PoseVertex and
 * LandmarkVertex, so the declarations below necessarily match the reference's; comments and everything behind the interface
 * are this repository's own.  User code written against the reference compiles unchanged:
 *
 *   Array<T,N>, Set<T>, UniquePtr<T>          reference types.h:36-43
 *   CameraParams {fx,fy,cx,cy,bf}             :51-62
 *   BaseEdge, Edge<DIM>, MonoEdge, StereoEdge :73-139   (scalar information, raw vertex pointers)
 *   EdgeType                                  :143-148
 *   PoseVertex, LandmarkVertex                :156-208  (q stored x,y,z,w; public `edges` sets; iP/iL)
 *   RobustKernelType                          :213-218
 *   BatchInfo, BatchStatistics, TimeProfile   :226-236
 *   VertexP, VertexL, Edge2D, Edge3D          :242-245
 *
 * Eigen: the reference needs <Eigen/Core> and <Eigen/Geometry>.  When Eigen is installed it is used;
 * otherwise the container-only stand-in under include/cuba_compat/ is picked up.
 */
#ifndef CUBA_B200_CUDA_BUNDLE_ADJUSTMENT_TYPES_H
#define CUBA_B200_CUDA_BUNDLE_ADJUSTMENT_TYPES_H

#include <map>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

#if defined(CUBA_FORCE_EIGEN_COMPAT)
#include "cuba_compat/Eigen/Core"
#include "cuba_compat/Eigen/Geometry"
#elif defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#include <Eigen/Geometry>
#else
#include "cuba_compat/Eigen/Core"
#include "cuba_compat/Eigen/Geometry"
#endif
#else
#include <Eigen/Core>
#include <Eigen/Geometry>
#endif

namespace cuba
{

template <class T, int N> using Array = Eigen::Matrix<T, N, 1>;
template <class T> using Set = std::unordered_set<T>;
template <class T> using UniquePtr = std::unique_ptr<T>;

/** Pinhole / rectified-stereo intrinsics of one view. bf = baseline * fx. */
struct CameraParams
{
	double fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
};

struct PoseVertex;
struct LandmarkVertex;

/** Type-erased edge: what the optimizer needs to walk the graph. */
struct BaseEdge
{
	virtual PoseVertex* poseVertex() const = 0;
	virtual LandmarkVertex* landmarkVertex() const = 0;
	virtual int dim() const = 0;
	virtual ~BaseEdge() {}
};

/** Reprojection edge with a DIM-dimensional pixel measurement and a scalar information value. */
template <int DIM>
struct Edge : BaseEdge
{
	using Measurement = Array<double, DIM>;
	using Information = double;

	Edge() : measurement(Measurement()), information(Information()), vertexP(nullptr), vertexL(nullptr) {}
	Edge(const Measurement& m, Information I, PoseVertex* vertexP, LandmarkVertex* vertexL)
		: measurement(m), information(I), vertexP(vertexP), vertexL(vertexL) {}

	PoseVertex* poseVertex() const override { return vertexP; }
	LandmarkVertex* landmarkVertex() const override { return vertexL; }
	int dim() const override { return DIM; }

	Measurement measurement;
	Information information;
	PoseVertex* vertexP;
	LandmarkVertex* vertexL;
};

using MonoEdge = Edge<2>;    // (u, v)
using StereoEdge = Edge<3>;  // (u_left, v, u_right)

enum class EdgeType { MONOCULAR = 0, STEREO = 1, COUNT = 2 };

/** SE(3) pose (world -> camera): Xc = R(q) Xw + t. */
struct PoseVertex
{
	using Quaternion = Eigen::Quaterniond;
	using Rotation = Quaternion;
	using Translation = Array<double, 3>;

	PoseVertex() : q(Rotation()), t(Translation()), fixed(false), id(-1), iP(-1) {}
	PoseVertex(int id, const Rotation& q, const Translation& t, const CameraParams& camera, bool fixed = false)
		: q(q), t(t), camera(camera), fixed(fixed), id(id), iP(-1) {}

	Rotation q;
	Translation t;
	CameraParams camera;
	bool fixed;
	int id;
	int iP;                 // internal index, written by initialize()
	Set<BaseEdge*> edges;   // maintained by add/remove edge
};

/** 3-D point in the world frame. */
struct LandmarkVertex
{
	using Point3D = Array<double, 3>;

	LandmarkVertex() : Xw(Point3D()), fixed(false), id(-1), iL(-1) {}
	LandmarkVertex(int id, const Point3D& Xw, bool fixed = false) : Xw(Xw), fixed(fixed), id(id), iL(-1) {}

	Point3D Xw;
	bool fixed;
	int id;
	int iL;                 // internal index, written by initialize()
	Set<BaseEdge*> edges;
};

enum class RobustKernelType { NONE = 0, HUBER = 1, TUKEY = 2 };

struct BatchInfo
{
	int iteration;
	double chi2;
};

using BatchStatistics = std::vector<BatchInfo>;
using TimeProfile = std::map<std::string, double>;

using VertexP = PoseVertex;
using VertexL = LandmarkVertex;
using Edge2D = MonoEdge;
using Edge3D = StereoEdge;

} // namespace cuba

#endif
