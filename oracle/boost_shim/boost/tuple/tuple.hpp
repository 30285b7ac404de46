// empty on purpose: only used by the reference under !NDEBUG (oracle builds use -DNDEBUG)
#pragma once
