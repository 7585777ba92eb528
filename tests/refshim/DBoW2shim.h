// stand-ins for Thirdparty/DBoW2/DBoW2/{BowVector.h:50-100, FeatureVector.h:21-45}
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId; typedef double WordValue; typedef unsigned int NodeId;
enum LNorm { L1, L2 };
class BowVector : public std::map<WordId, WordValue> { public: void addWeight(WordId id, WordValue v); void addIfNotExist(WordId id, WordValue v); void normalize(LNorm norm_type); };
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > { public: void addFeature(NodeId id, unsigned int i_feature); };
}
