// TEST INFRASTRUCTURE ONLY: src/Frame.cc includes <include/LocalMapping.h> but uses nothing from it; the real header drags in the whole back end.
#pragma once
