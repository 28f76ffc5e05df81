// tests/host/cluster.h -- TEST STAND-IN for the reference's include/cluster.h (not present on the GPU box):
// only the POD the ORBextractor interface names (/root/reference/include/cluster.h:22-31) and the
// using-directives that header injects (:17-18), which the reference's ORBextractor.h relies on.
#ifndef CLUSTER_H
#define CLUSTER_H
#include <vector>
#include <opencv2/core/core.hpp>
using namespace cv;
using namespace std;
namespace ORB_SLAM2 {
struct center { int x, y, L, A, B, D, label, id; };
}
#endif
