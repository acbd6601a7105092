// ORACLE — TEST INFRASTRUCTURE ONLY.
// ref_knn.cpp: builds the REFERENCE's own vendored nanoflann (thirdparty/nanoflann.hpp v1.4.2 and
// thirdparty/KDTreeVectorOfVectorsAdaptor.h, included from where they lie under /root/reference at build time — never
// copied) behind the call sequence of TreeBucket::knnSearch (src/cont2/contour_db.cpp:381-403): KD-tree with leaf size
// 10 over RET_KEY_DIM = 10 float keys (contour_db.h:115-116), metric_L2, a KNN result set whose worst distance is
// initialised to max_dist_sq (MyKNNResSet, contour_db.h:32-52 — restated here in 8 lines because contour_db.h itself
// needs Eigen/OpenCV/Ceres), SearchParams(10).  Output goes to oracle/_ref/libref_knn.so (git-ignored, travels to the
// GPU box).  Used by tests to validate the exhaustive-scan kNN of oracle/c2o_query.hpp.
#include <array>
#include <cstdint>  // nanoflann.hpp 1.4.2 relies on a transitive <cstdint> that g++ 13 no longer provides
#include <vector>

#include <nanoflann.hpp>
#include <KDTreeVectorOfVectorsAdaptor.h>

typedef std::array<float, 10> RKey;
typedef std::vector<RKey> my_vector_of_vectors_t;
typedef KDTreeVectorOfVectorsAdaptor<my_vector_of_vectors_t, float> my_kd_tree_t;

template <typename D, typename I = size_t, typename C = size_t>
class MyKNNResSet : public nanoflann::KNNResultSet<D, I, C> {
 public:
  explicit MyKNNResSet(C capacity_) : nanoflann::KNNResultSet<D, I, C>(capacity_) {}
  void init(I *indices_, D *dists_, D max_dist_metric) {
    this->indices = indices_;
    this->dists = dists_;
    this->count = 0;
    if (this->capacity) this->dists[this->capacity - 1] = max_dist_metric;
  }
};

struct RefTree {
  my_vector_of_vectors_t data;
  my_kd_tree_t *tree = nullptr;
  ~RefTree() { delete tree; }
};

extern "C" {
void *ref_knn_build(const float *keys, int n) {
  RefTree *t = new RefTree;
  t->data.resize(n);
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 10; ++d) t->data[i][d] = keys[i * 10 + d];
  if (n > 0) t->tree = new my_kd_tree_t(10, t->data, 10);
  return t;
}
void ref_knn_free(void *h) { delete (RefTree *) h; }
// mirrors TreeBucket::knnSearch; out_idx/out_dist have num_res entries; unfilled distance slots keep 1e6
void ref_knn_search(void *h, const float *q, int num_res, float max_dist_sq, int64_t *out_idx, float *out_dist) {
  RefTree *t = (RefTree *) h;
  std::vector<size_t> idx(num_res, 0);
  for (int i = 0; i < num_res; ++i) out_dist[i] = 1e6f;
  if (t->tree) {
    MyKNNResSet<float> resultSet(num_res);
    resultSet.init(&idx[0], out_dist, max_dist_sq);
    t->tree->index->findNeighbors(resultSet, q, nanoflann::SearchParams(10));
  }
  for (int i = 0; i < num_res; ++i) out_idx[i] = (int64_t) idx[i];
}
}
