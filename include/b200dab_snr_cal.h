/* b200dab_snr_cal.h -- this path's decision-directed SNR estimate mapped onto the scale of the reference core's snr10
 * (dabsdr.h:153).  MEASURED from libdabsdr 4.0.1 as a black box over AWGN files from 2 to 45 dB
 * (tests/golden/make_snr_cal.py): {raw estimate dB, reference dB}, ascending; interpolate linearly, extrapolate with slope 1. */
#ifndef B200DAB_SNR_CAL_H
#define B200DAB_SNR_CAL_H
#define B200DAB_SNR_CAL_N 35
static const float B200DAB_SNR_CAL[B200DAB_SNR_CAL_N][2] = {
    {2.00f, 4.00f}, {2.51f, 4.20f}, {3.09f, 4.40f}, {3.76f, 4.70f}, {4.50f, 5.20f}, {5.31f, 5.80f},
    {6.20f, 6.70f}, {7.14f, 7.60f}, {8.13f, 8.70f}, {9.12f, 9.80f}, {10.14f, 10.90f}, {11.17f, 12.00f},
    {12.18f, 13.20f}, {13.20f, 14.30f}, {14.20f, 15.40f}, {15.21f, 16.40f}, {16.21f, 17.50f}, {17.22f, 18.50f},
    {18.22f, 19.60f}, {19.23f, 20.60f}, {20.24f, 21.60f}, {21.24f, 22.60f}, {22.24f, 23.60f}, {23.24f, 24.50f},
    {24.24f, 25.50f}, {25.24f, 26.40f}, {26.24f, 27.40f}, {27.24f, 28.30f}, {28.24f, 29.20f}, {30.24f, 31.10f},
    {32.25f, 33.00f}, {34.24f, 35.00f}, {36.25f, 37.00f}, {38.24f, 38.90f}, {43.25f, 43.80f},
};
static inline float b200dab_snr_calibrated(float raw_db)
{
    const int n = B200DAB_SNR_CAL_N;
    if (raw_db <= B200DAB_SNR_CAL[0][0]) return raw_db + (B200DAB_SNR_CAL[0][1] - B200DAB_SNR_CAL[0][0]);
    for (int i = 1; i < n; i++)
        if (raw_db <= B200DAB_SNR_CAL[i][0]) {
            const float t = (raw_db - B200DAB_SNR_CAL[i - 1][0]) / (B200DAB_SNR_CAL[i][0] - B200DAB_SNR_CAL[i - 1][0]);
            return B200DAB_SNR_CAL[i - 1][1] + t * (B200DAB_SNR_CAL[i][1] - B200DAB_SNR_CAL[i - 1][1]);
        }
    return raw_db + (B200DAB_SNR_CAL[n - 1][1] - B200DAB_SNR_CAL[n - 1][0]);
}
#endif
