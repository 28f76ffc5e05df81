// tests/host/KeyFrame.h -- TEST STAND-IN (the on-path methods never touch a KeyFrame)
#ifndef KEYFRAME_H
#define KEYFRAME_H
namespace ORB_SLAM2 { class KeyFrame; }
#endif
