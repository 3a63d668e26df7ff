/*
 * dabsdr_abi.h -- binary interface of the "dabsdr" receiver core as consumed by AbracaDABra's RadioControl.
 *
 * This is this repository's own declaration of the C ABI that `libb200dabsdr.so` (SONAME libdabsdr.so.4)
 * exports, so that the library is a drop-in for the closed `libdabsdr.so.4.0.1` the reference links
 * (reference interface: lib/linux_x86_64/dabsdr.h:37-436; sole caller: src/radiocontrol.cpp, wrappers at
 * src/radiocontrol.h:539-559).  Struct layouts, enum values and symbol names must stay identical to that
 * interface -- tests/test_abi.py checks sizeof/offsetof of every type against values recorded from the
 * reference header, and tests/test_shim.py drives both libraries with the same harness binary.
 *
 * Sections: (1) handle/version, (2) payload callbacks, (3) notifications, (4) entry points.
 */
#ifndef B200_DABSDR_ABI_H
#define B200_DABSDR_ABI_H

#include <stdint.h>
#include <stdlib.h>

#if defined(__GNUC__)
#define DABSDR_API __attribute__((visibility("default")))
#else
#define DABSDR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) handle and version: dabsdr.h:37-45 */
typedef struct dabsdr_s *dabsdrHandle_t;
typedef struct { uint8_t major, minor, patch, flags; } dabsdrVersion_t;

/* ---- (2) payload callbacks */
/* Audio frame header byte (dabsdr.h:47-60): for DAB+ it is superframe byte 2 with bit 7 = concealment flag;
 * for MP2 the same byte carries the DRC value. */
typedef union {
    uint8_t raw;
    uint8_t mp2DRC;
    struct {
        uint8_t mpeg_surr_cfg : 3;
        uint8_t ps_flag : 1;
        uint8_t aac_channel_mode : 1;
        uint8_t sbr_flag : 1;
        uint8_t dac_rate : 1;
        uint8_t conceal : 1;
    } bits;
} dabsdrAudioFrameHeader_t;

/* decoder instances (dabsdr.h:62-67) */
typedef enum dabsdrDecoderId_e {
    DABSDR_ID_DATA = -1,
    DABSDR_ID_AUDIO_PRIMARY = 0,
    DABSDR_ID_AUDIO_SECONDARY = 1,
} dabsdrDecoderId_t;

/* one DAB+ access unit or MP2 frame (dabsdr.h:71-78); pAuData is valid during the callback only */
typedef struct {
    dabsdrDecoderId_t id;
    uint8_t ASCTy;
    dabsdrAudioFrameHeader_t header;
    uint16_t auLen;
    const uint8_t *pAuData;
} dabsdrAudioCBData_t;

/* dynamic label segment (dabsdr.h:81-86) */
typedef struct {
    dabsdrDecoderId_t id;
    uint16_t len;
    const uint8_t *pData;
} dabsdrDynamicLabelCBData_t;

/* MSC data group (dabsdr.h:89-96) */
typedef struct {
    dabsdrDecoderId_t id;
    uint16_t SCId;
    uint16_t userAppType;
    uint16_t dgLen;
    const uint8_t *pDgData;
} dabsdrDataGroupCBData_t;

/* ---- (3) notifications (dabsdr.h:99-135) */
typedef enum dabsdrNotificationId_e {
    DABSDR_NID_SYNC_STATUS = 1,
    DABSDR_NID_TUNE = 2,
    DABSDR_NID_ENSEMBLE_INFO = 3,
    DABSDR_NID_SERVICE_LIST = 4,
    DABSDR_NID_SERVICE_COMPONENT_LIST = 5,
    DABSDR_NID_USER_APP_UPDATE = 6,
    DABSDR_NID_USER_APP_LIST = 7,
    DABSDR_NID_SERVICE_SELECTION = 8,
    DABSDR_NID_SERVICE_STOP = 9,
    DABSDR_NID_PERIODIC = 10,
    DABSDR_NID_XPAD_APP_START_STOP = 11,
    DABSDR_NID_RECONFIGURATION = 12,
    DABSDR_NID_RESET = 13,
    DABSDR_NID_ANNOUNCEMENT_SUPPORT = 14,
    DABSDR_NID_ANNOUNCEMENT_SWITCHING = 15,
    DABSDR_NID_PTY = 16,
    DABSDR_NID_TII = 17,
} dabsdrNotificationId_t;

typedef enum dabsdrNotificationStatus_e {
    DABSDR_NSTAT_SUCCESS = 0,
    DABSDR_NSTAT_GENERIC_ERROR = 1,
    DABSDR_NSTAT_SERVICE_NOT_FOUND = 2,
    DABSDR_NSTAT_SERVICE_NOT_READY = 3,
    DABSDR_NSTAT_SERVICE_NOT_SUPPORTED = 4,
} dabsdrNotificationStatus_t;

typedef struct {
    dabsdrNotificationId_t nid;
    dabsdrNotificationStatus_t status;
    uint16_t len;             /* sizeof the struct behind pData */
    const void *pData;        /* valid during the callback only; may be NULL (first PERIODIC) */
} dabsdrNotificationCBData_t;

/* sync level (dabsdr.h:138-143) */
typedef enum {
    DABSDR_SYNC_LEVEL_NO_SYNC = 0,
    DABSDR_SYNC_LEVEL_ON_NULL = 1,
    DABSDR_SYNC_LEVEL_FIC = 3,
} dabsdrSyncLevel_t;

typedef struct {
    dabsdrSyncLevel_t syncLevel;
    int16_t snr10;            /* 10 x SNR[dB] */
} dabsdrNtfSyncStatus_t;

/* statistics every 2^period frames (dabsdr.h:151-166); 32 bytes */
typedef struct {
    dabsdrSyncLevel_t syncLevel;
    int16_t snr10;
    int32_t freqOffset;             /* Hz x 10 */
    uint32_t dateHoursMinutes;
    uint16_t secMsec;
    uint16_t fibErrorCntr;
    uint8_t mscCrcOkCntr;
    uint8_t mscCrcErrorCntr;
    uint16_t audioServiceBytes;
    uint16_t padBytes;
    uint16_t rsUncorrectableCntr;
    uint16_t rsBitErrors;
    uint16_t rsBytes;
} dabsdrNtfPeriodic_t;

#define DAB_LABEL_MAX_LENGTH (16)
typedef struct {
    char str[DAB_LABEL_MAX_LENGTH + 1];
    uint16_t charField;
    uint8_t charset;
} dabsdrLabel_t;

/* service list item (dabsdr.h:177-191) */
typedef struct {
    uint32_t sid;
    dabsdrLabel_t label;
    struct { uint8_t s; uint8_t d; } pty;
    uint8_t CAId;
} dabsdrServiceListItem_t;

/* service component (dabsdr.h:194-243) */
typedef struct {
    uint8_t SCIdS;
    uint8_t SubChId;
    int16_t SubChAddr;              /* CU address 0..863, < 0 = not yet known */
    uint16_t SubChSize;
    uint8_t protectionLevel;
    union { uint8_t uepIdx; uint8_t fecScheme; };
    uint8_t ps;
    uint8_t lang;
    uint8_t CAflag;
    dabsdrLabel_t label;
    uint8_t numUserApps;
    uint8_t TMId;                   /* 0 stream audio, 1 stream data, 3 packet data */
    union {
        struct { uint8_t ASCTy; uint16_t bitRate; } streamAudio;
        struct { uint8_t DSCTy; uint16_t bitRate; } streamData;
        struct { uint8_t DSCTy; uint16_t SCId; uint8_t DGflag; int16_t packetAddress; } packetData;
    };
} dabsdrServiceCompListItem_t;

typedef struct {
    uint16_t type;
    dabsdrLabel_t label;
    uint8_t dataLen;
    uint8_t data[23];
} dabsdrUserAppListItem_t;

/* list notifications carry a count and an accessor to be called inside the callback (dabsdr.h:260-264) */
typedef struct {
    uint8_t numServices;
    int (*getServiceListItem)(dabsdrHandle_t handle, uint8_t idx, dabsdrServiceListItem_t *pServiceListItem);
} dabsdrNtfServiceList_t;

typedef struct {
    uint32_t SId;
    uint8_t numServiceComponents;
    int (*getServiceComponentListItem)(dabsdrHandle_t handle, uint8_t scIdx, dabsdrServiceCompListItem_t *pServiceCompListItem);
} dabsdrNtfServiceComponentList_t;

typedef struct {
    uint32_t SId;
    uint16_t ASu;
    uint8_t numClusterIds;
    uint8_t clusterIds[7];
} dabsdrNtfAnnouncementSupport_t;

typedef struct { uint32_t SId; uint8_t SCIdS; } dabsdrNtfUserAppUpdate_t;

typedef struct {
    uint32_t SId;
    uint8_t SCIdS;
    uint8_t numUserApps;
    int (*getUserAppListItem)(dabsdrHandle_t handle, uint8_t uaIdx, dabsdrUserAppListItem_t *pUserAppListItem);
} dabsdrNtfUserAppList_t;

/* ensemble (dabsdr.h:301-318): ueid = ECC << 16 | EId */
typedef struct {
    uint32_t frequency;
    union {
        uint32_t ueid;
        struct { uint32_t eid : 16; uint32_t ecc : 8; };
    };
    int8_t LTO;
    uint8_t intTable;
    uint8_t alarm;
    dabsdrLabel_t label;
} dabsdrNtfEnsemble_t;

typedef struct {
    uint32_t SId;
    uint8_t SCIdS;
    dabsdrDecoderId_t id;
} dabsdrNtfServiceSelection_t;
typedef dabsdrNtfServiceSelection_t dabsdrNtfServiceStop_t;

typedef enum dabsdrNtfResetFlags_e { DABSDR_RESET_INIT = 0, DABSDR_RESET_NEW_EID = 1 } dabsdrNtfResetFlags_t;

typedef struct { uint8_t appType; int8_t start; } dabsdrNtfXpadAppStartStop_t;
typedef struct { uint8_t clusterId; uint8_t subChId; uint16_t ASwFlags; } dabsdrAsw_t;
typedef struct { dabsdrAsw_t asw[8]; } dabsdrNtfAnnouncementSwitching_t;
typedef struct { uint32_t SId; uint8_t s; uint8_t d; } dabsdrNtfPTy_t;

/* TII (dabsdr.h:359-384) */
typedef enum dabsdrTiiMode_e { DABSDR_TII_MODE_CONSERVATIVE = 0, DABSDR_TII_MODE_DEFAULT = 1, DABSDR_TII_NUM_MODES } dabsdrTiiMode_t;
typedef enum dabsdrSpectrum_e { DABSDR_SPECT_SIGNAL = 0, DABSDR_SPECT_NULL = 1, DABSDR_SPECT_TII = 2, DABSDR_SPECT_NUM_TYPES } dabsdrSpectrum_t;
typedef struct { uint8_t main; uint8_t sub; float level; } dabsdrTii_t;
typedef struct {
    uint8_t numIds;
    dabsdrTii_t id[24];
    /* writes 384 floats (2 x 192 carrier pairs); the caller allocates 384: src/radiocontrol.h:280 */
    int (*getSpectrumTii)(dabsdrHandle_t handle, float buffer[192]);
} dabsdrNtfTii_t;

/* ---- (4) entry points (dabsdr.h:387-429) */
typedef void (*dabsdrInputFunc_t)(float[], uint16_t);
typedef void (*dabsdrAudioCBFunc_t)(dabsdrAudioCBData_t *p, void *ctx);
typedef void (*dabsdrDynamicLabelCBFunc_t)(dabsdrDynamicLabelCBData_t *p, void *ctx);
typedef void (*dabsdrDataGroupCBFunc_t)(dabsdrDataGroupCBData_t *p, void *ctx);
typedef void (*dabsdrSpectrumCBFunc_t)(const float *p, dabsdrSpectrum_t type, void *ctx);
typedef void (*dabsdrNotificationCBFunc_t)(dabsdrNotificationCBData_t *p, void *ctx);

DABSDR_API uint8_t dabsdrInit(dabsdrHandle_t *handle);       /* 0 on success: src/radiocontrol.cpp:83 */
DABSDR_API void dabsdr(dabsdrHandle_t handle);               /* starts the worker thread and returns: :93 */
DABSDR_API void dabsdrDeinit(dabsdrHandle_t *handle);        /* after dabsdrRequest_Exit: :77, :420-426 */
DABSDR_API void dabsdrGetVersion(dabsdrVersion_t *version);

DABSDR_API void dabsdrRegisterInputFcn(dabsdrHandle_t handle, dabsdrInputFunc_t fcn);       /* getSamples */
DABSDR_API void dabsdrRegisterDummyInputFcn(dabsdrHandle_t handle, dabsdrInputFunc_t fcn);  /* skipSamples */
DABSDR_API void dabsdrRegisterAudioCb(dabsdrHandle_t handle, dabsdrAudioCBFunc_t fcn, void *ctx);
DABSDR_API void dabsdrRegisterDynamicLabelCb(dabsdrHandle_t handle, dabsdrDynamicLabelCBFunc_t fcn, void *ctx);
DABSDR_API void dabsdrRegisterDataGroupCb(dabsdrHandle_t handle, dabsdrDataGroupCBFunc_t fcn, void *ctx);
DABSDR_API void dabsdrRegisterSignalSpectrumCb(dabsdrHandle_t handle, dabsdrSpectrumCBFunc_t fcn, void *ctx);
DABSDR_API void dabsdrRegisterNotificationCb(dabsdrHandle_t handle, dabsdrNotificationCBFunc_t fcn, void *ctx);

DABSDR_API void dabsdrRequest_Tune(dabsdrHandle_t handle, uint32_t frequency);
DABSDR_API void dabsdrRequest_GetEnsemble(dabsdrHandle_t handle);
DABSDR_API void dabsdrRequest_GetServiceList(dabsdrHandle_t handle);
DABSDR_API void dabsdrRequest_GetServiceComponents(dabsdrHandle_t handle, uint32_t SId);
DABSDR_API void dabsdrRequest_GetUserAppList(dabsdrHandle_t handle, uint32_t SId, uint8_t SCIdS);
DABSDR_API void dabsdrRequest_GetAnnouncementSupport(dabsdrHandle_t handle, uint32_t SId);
DABSDR_API void dabsdrRequest_ServiceSelection(dabsdrHandle_t handle, uint32_t SId, uint8_t SCIdS, dabsdrDecoderId_t id);
DABSDR_API void dabsdrRequest_ServiceStop(dabsdrHandle_t handle, uint32_t SId, uint8_t SCIdS, dabsdrDecoderId_t id);
DABSDR_API void dabsdrRequest_XPadAppStart(dabsdrHandle_t handle, uint8_t appType, int8_t startRequest, dabsdrDecoderId_t id);
DABSDR_API void dabsdrRequest_SetPeriodicNotify(dabsdrHandle_t handle, uint8_t period, uint32_t cfg);
DABSDR_API void dabsdrRequest_SetTII(dabsdrHandle_t handle, uint8_t ena, dabsdrTiiMode_t mode);
DABSDR_API void dabsdrRequest_SignalSpectrum(dabsdrHandle_t handle, uint8_t ena);
DABSDR_API void dabsdrRequest_Exit(dabsdrHandle_t handle);

#ifdef __cplusplus
}
#endif
#endif /* B200_DABSDR_ABI_H */
