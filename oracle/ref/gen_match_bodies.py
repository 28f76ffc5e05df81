#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Build step of oracle/_ref: copies the bodies of the hot-path matcher functions out of
the reference's own sources, verbatim, into a GENERATED include file under oracle/_ref/gen/ (git-ignored build
output -- nothing is copied into the repository).  The functions are located by signature and brace matching:

  src/ORBmatcher.cc : TH_HIGH/TH_LOW/HISTO_LENGTH + ctor, SearchByProjection(Frame&, vector<MapPoint*>&, th),
                      RadiusByViewingCos, SearchForInitialization, SearchByProjection(Frame&, const Frame&, th, bMono),
                      ComputeThreeMaxima, DescriptorDistance
                      SearchByBoW(KeyFrame*, Frame&, ...), SearchByBoW(KeyFrame*, KeyFrame*, ...),
                      SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)
  src/KeyFrame.cc   : GetFeaturesInArea, IsInImage
  src/MapPoint.cc   : ComputeDistinctiveDescriptors
  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h : transform(features, BowVector, FeatureVector, levelsup), transform(feature, ...)
  src/cluster.cc    : clustering, updateCenter, initilizeCenters, fituneCenter, SLIC, initClusterAssment, loadDataSet, distEclud, kmeans
  src/Frame.cc      : AssignFeaturesToGrid, GetFeaturesInArea, PosInGrid, ComputeStereoMatches,
                      UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD

usage: gen_match_bodies.py <reference root> <out dir>
"""
import os, re, sys

ref, out = sys.argv[1], sys.argv[2]


def extract(path, starts):
    src = open(os.path.join(ref, path), encoding="utf-8", errors="replace").read()
    chunks = []
    for sig in starts:
        i = src.find(sig)
        assert i >= 0, (path, sig)
        assert src.find(sig, i + 1) < 0, ("ambiguous", sig)
        j = src.index("{", i)
        depth, k = 0, j
        in_line_c = in_block_c = False
        in_str = None
        while True:
            c = src[k]; n2 = src[k:k + 2]
            if in_line_c:
                if c == "\n": in_line_c = False
            elif in_block_c:
                if n2 == "*/": in_block_c = False; k += 1
            elif in_str:
                if c == "\\": k += 1
                elif c == in_str: in_str = None
            elif n2 == "//": in_line_c = True; k += 1
            elif n2 == "/*": in_block_c = True; k += 1
            elif c in "\"'": in_str = c
            elif c == "{": depth += 1
            elif c == "}":
                depth -= 1
                if depth == 0: break
            k += 1
        line0 = src.count("\n", 0, i) + 1
        chunks.append("// ---- %s:%d ----\n#line %d \"%s\"\n%s\n" % (path, line0, line0, os.path.join(ref, path), src[i:k + 1]))
    return "\n".join(chunks)


m = extract("src/ORBmatcher.cc", [
    "const int ORBmatcher::TH_HIGH = 100;\nconst int ORBmatcher::TH_LOW = 50;\nconst int ORBmatcher::HISTO_LENGTH = 30;",
    "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)",
    "float ORBmatcher::RadiusByViewingCos(const float &viewCos)",
    "int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize)",
    "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)",
    "void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)",
    "int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)",
    "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th , const int ORBdist)",
    "int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)",
    "int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12,\n                             const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)",
    "bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint &kp1,const cv::KeyPoint &kp2,const cv::Mat &F12,const KeyFrame* pKF2)",
    "int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,\n                                       vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)",
    "int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)",
    "int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)",
    "int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F, vector<MapPoint*> &vpMapPointMatches)",
    "int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12)",
])
f = extract("src/Frame.cc", [
    "void Frame::AssignFeaturesToGrid()",
    "vector<size_t> Frame::GetFeaturesInArea(const float &x, const float  &y, const float  &r, const int minLevel, const int maxLevel) const",
    "bool Frame::PosInGrid(",
    "void Frame::ComputeStereoMatches()",
    "void Frame::UndistortKeyPoints()",
    "void Frame::ComputeImageBounds(const cv::Mat &imLeft)",
    "void Frame::ComputeStereoFromRGBD(const cv::Mat &imDepth)",
])
kf = extract("src/KeyFrame.cc", [
    "vector<size_t> KeyFrame::GetFeaturesInArea(const float &x, const float &y, const float &r) const",
    "bool KeyFrame::IsInImage(const float &x, const float &y) const",
])
mp = extract("src/MapPoint.cc", ["void MapPoint::ComputeDistinctiveDescriptors()"])
# DBoW2 (vendored in the reference tree): the two transform() members of the vocabulary template, for oracle/ref/ref_bow_capi.cpp
b = extract("Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h", [
    "void TemplatedVocabulary<TDescriptor,F>::transform(\n  const std::vector<TDescriptor>& features,\n  BowVector &v, FeatureVector &fv, int levelsup) const",
    "void TemplatedVocabulary<TDescriptor,F>::transform(const TDescriptor &feature, \n  WordId &word_id, WordValue &weight, NodeId *nid, int levelsup) const",
])
b = b.replace("void TemplatedVocabulary<TDescriptor,F>::transform(", "template<class TDescriptor, class F>\nvoid TemplatedVocabulary<TDescriptor,F>::transform(")
# SLIC stage of `cluster` (the k-means that follows seeds itself with rand(): canonical-seed contract, see oracle/ref/ref_slic_capi.cpp)
sl = extract("src/cluster.cc", [
    "int cluster::clustering(const cv::Mat &imageLAB,const cv::Mat &DepthImage, cv::Mat &DisMask, cv::Mat &labelMask,",
    "int cluster::updateCenter(cv::Mat &imageLAB, cv::Mat &labelMask,cv::Mat const &Depth,std::vector<center> &centers, int len)",
    "int cluster::initilizeCenters(cv::Mat &imageLAB,cv::Mat const &imagedepth, std::vector<center> &centers, int len)",
    "int cluster::fituneCenter(cv::Mat &imageLAB, cv::Mat &sobelGradient, std::vector<center> &centers)",
    "int cluster::SLIC(cv::Mat const &image,cv::Mat const &image_D, cv::Mat &resultLabel, std::vector<center> &centers, int len, int m)",
    "void cluster::initClusterAssment()",
    "void cluster::loadDataSet(vector<center> &centers)",
    "double cluster::distEclud(center &v1 ,center &v2)",
    "void cluster::kmeans()",
])
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "ref_slic_bodies.inc"), "w").write("// GENERATED from the reference sources by oracle/ref/gen_match_bodies.py -- do not commit\n" + sl)
open(os.path.join(out, "ref_bow_bodies.inc"), "w").write("// GENERATED from the reference sources by oracle/ref/gen_match_bodies.py -- do not commit\n" + b)
# the first ORBmatcher chunk (constants) ends at the ctor's closing brace because the ctor follows immediately
open(os.path.join(out, "ref_match_bodies.inc"), "w").write("// GENERATED from the reference sources by oracle/ref/gen_match_bodies.py -- do not commit\n" + m + "\n" + f + "\n" + kf + "\n" + mp)
print("generated", os.path.join(out, "ref_match_bodies.inc"))
