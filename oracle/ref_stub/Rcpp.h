/* ref_stub/Rcpp.h — TEST INFRASTRUCTURE (oracle). See RcppArmadillo.h in this directory. */
#include "RcppArmadillo.h"
