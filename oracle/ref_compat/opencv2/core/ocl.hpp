/* oracle/ref_compat: empty stand-in (TEST INFRASTRUCTURE) -- see ../opencv.hpp */
