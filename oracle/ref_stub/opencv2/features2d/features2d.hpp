// stand-in for <opencv2/features2d/features2d.hpp> (nothing of it is used by the files compiled into oracle/_ref)
#include "../core/core.hpp"
