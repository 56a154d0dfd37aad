// TEST INFRASTRUCTURE (oracle build only): empty stand-in. OpenCV C++ headers are absent from this image and
// include/torchlib/utils.h names no cv:: symbol on the hot path.
#pragma once
