// tests/host/DBoW2_standin.h -- TEST STAND-IN for Thirdparty/DBoW2/DBoW2/{BowVector,FeatureVector}.h and include/ORBVocabulary.h:
// the container types exactly as DBoW2 declares them (BowVector.h:21-23, 54-56; FeatureVector.h:20-22), the vocabulary opaque.
#ifndef DBOW2_STANDIN_H
#define DBOW2_STANDIN_H
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}
namespace ORB_SLAM2 { class ORBVocabulary {}; }              // include/ORBVocabulary.h:40-43 (a DBoW2 template instance there)
#endif
