// OpenCV-free shim header (TEST INFRASTRUCTURE ONLY): forwards to oracle/cvlite/cvlite.hpp so that
// the reference's own sources compile unmodified into oracle/_ref/ (see oracle/ref/Makefile).
#include "../../cvlite/cvlite.hpp"
