/*
 * oracle/ref_build/stubs.c - TEST INFRASTRUCTURE.
 * Empty definitions of the 18 symbols provided by the four reference sources
 * that genuinely need HDF5 and are therefore left out of the oracle build
 * (line_of_sight.c, lightcone/lightcone.c, lightcone/lightcone_particle_io.c,
 * neutrino/Default/neutrino_response.c). None is reachable from the hydro
 * hot path; they only satisfy the linker for engine.c / restart.c.
 */
#define STUB(name) \
  void name(void) {}
STUB(do_line_of_sight)
STUB(lightcone_clean)
STUB(lightcone_dump_completed_shells)
STUB(lightcone_flush_map_updates)
STUB(lightcone_flush_particle_buffers)
STUB(lightcone_init)
STUB(lightcone_memory_use)
STUB(lightcone_prepare_for_step)
STUB(lightcone_struct_dump)
STUB(lightcone_struct_restore)
STUB(lightcone_trigger_map_update)
STUB(lightcone_write_index)
STUB(los_io_output_check)
STUB(los_struct_dump)
STUB(los_struct_restore)
STUB(neutrino_response_struct_dump)
STUB(neutrino_response_struct_restore)
int io_is_double_precision(int field) {
  (void)field;
  return 0;
}
