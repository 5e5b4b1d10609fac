/* Empty stub of <gsl/gsl_const_num.h> (globals.h:23). TEST INFRASTRUCTURE ONLY. */
