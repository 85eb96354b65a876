// opencv2/imgcodecs.hpp stand-in (test infrastructure): the whole slice of cv:: the reference uses lives in cvshim.hpp
#pragma once
#include "../cvshim.hpp"
