double orc_pert_a = 1.0, orc_pert_d = 0.0;
int orc_pert_mode = 0;
void orc_set_perturb(double a, double d, int mode) { orc_pert_a = a, orc_pert_d = d, orc_pert_mode = mode; }
int orc_trace_on = 0;
void orc_set_trace(int on) { orc_trace_on = on; }
