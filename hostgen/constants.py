"""Unit constants in the reference's "Mpc units" (hbar = c = k_B = 1, lengths in Mpc).

Restates the constants the reference obtains from Unitful/UnitfulCosmo/PhysicalConstants
(CODATA2018), which are not vendored under the reference tree:
  km_s_Mpc_100, G_natural, mass_natural      src/Bolt.jl:49-51
  m_H, sigma_T (m_e, alpha unused here)      src/ionization/ionization.jl:34-39
  H0_natural_unit_conversion, Kelvin_...     src/ionization/recfast.jl:4-5
"""
import math

# SI (CODATA 2018, exact where defined)
C_SI = 299792458.0
HBAR_SI = 6.62607015e-34 / (2.0 * math.pi)
KB_SI = 1.380649e-23
EV_SI = 1.602176634e-19
G_SI = 6.67430e-11
AU_SI = 149597870700.0
PC_SI = AU_SI * 648000.0 / math.pi
MPC_SI = 1.0e6 * PC_SI
PROTON_MASS_SI = 1.67262192369e-27
THOMSON_SI = 6.6524587321e-29

# 100 km/s/Mpc in Mpc^-1
km_s_Mpc_100 = 100.0e3 / C_SI
# Newton's constant in Mpc^2  (G hbar / c^3 = Planck length squared)
G_natural = G_SI * HBAR_SI / C_SI**3 / MPC_SI**2
# 1 eV in Mpc^-1
mass_natural = EV_SI / (HBAR_SI * C_SI) * MPC_SI
# proton mass in Mpc^-1, Thomson cross-section in Mpc^2
m_H = PROTON_MASS_SI * C_SI / HBAR_SI * MPC_SI
sigma_T = THOMSON_SI / MPC_SI**2
# one natural time unit (1 Mpc / c) in seconds
H0_natural_unit_conversion = MPC_SI / C_SI
# one natural temperature unit (1 Mpc^-1) in Kelvin
Kelvin_natural_unit_conversion = HBAR_SI * C_SI / (KB_SI * MPC_SI)

ZETA3 = 1.2020569  # src/background.jl:3
