// hmesh/hmesh.h:7 includes <uuid/uuid.h> (libuuid, absent from the image) without using it on this path.
#pragma once
